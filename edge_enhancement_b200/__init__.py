"""edge_enhancement_b200 -- B200-native (sm_100a) implementation of the data-parallel hot path of
Aiqz/Edge-Enhancement: the Canny-style edge-enhancement transform (forward + adjoint, fused with
the blend) and the PGD/FGSM inner-loop updates, behind the reference's own Python surface.

    from edge_enhancement_b200 import core, attacks        # drop-ins for utils.core / utils.attacks
    edge_enhancement_b200.install()                        # or: make `import utils.core` resolve here

``edge-enhancement_b200`` at the repo root is a symlink to this directory (the project's name has a hyphen,
which Python cannot import).
"""
import sys as _sys

from . import _build, _lib            # noqa: F401
from . import functional              # noqa: F401
from . import core                    # noqa: F401
from . import attacks                 # noqa: F401
from .core import (CannyFilter, CannyFilter_BPDA, CannyFilter_step125_1, EdgeEnhance,  # noqa: F401
                   HighFreqSuppress, Add_Square, edge_enhance)
from .attacks import PGD, FGSM, Trades  # noqa: F401

__version__ = "0.1.0"


def build(force=False, verbose=False):
    """Compile libedge_b200.so for sm_100a if it is missing or stale; returns its path."""
    return _build.build(force=force, verbose=verbose)


def install(package="utils", shims=False):
    """Make ``from utils.core import CannyFilter`` / ``from utils.attacks import PGD`` (the imports
    at the top of every reference experiment script) resolve to this package's drop-ins.  Other
    ``utils.*`` modules (helper, data_loader, ...) keep resolving to the reference's own files.
    shims=True additionally registers stand-ins for the scripts' small third-party imports
    (easydict, managpu, autoattack) when those are not installed (compat.py)."""
    import importlib
    if shims:
        from . import compat
        compat.install_shims()
    try:
        pkg = importlib.import_module(package)
    except ImportError:
        import types
        pkg = types.ModuleType(package)
        pkg.__path__ = []
        _sys.modules[package] = pkg
    _sys.modules[package + ".core"] = core
    _sys.modules[package + ".attacks"] = attacks
    pkg.core = core
    pkg.attacks = attacks
    return pkg
