"""Importable alias of the ``edge-enhancement_b200/`` directory (a hyphen cannot be imported)."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "edge-enhancement_b200")
__path__[:] = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
