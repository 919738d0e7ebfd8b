"""Import the UNMODIFIED reference modules (utils.core / utils.attacks) from /root/reference.

TEST INFRASTRUCTURE ONLY, and only usable in the build container: /root/reference does not
exist on the GPU box, so nothing that runs there (pytest -m gpu, smoke(), bench.py) may call
this.  It is used by oracle/make_golden.py to produce tests/golden/*.npz and by the
container-only oracle-vs-live-reference tests (skipped when the reference is absent).

The only obstacle on torch >= 2.0 is ``from torch._six import builtins`` in the reference's
vendored utils/_jit_internal.py:9; a 3-line module shim fixes it without touching the
reference.
"""
import builtins
import contextlib
import io
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("EE_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "utils", "core.py"))


def load():
    """Returns (ref_core, ref_attacks) -- the reference's own modules."""
    if not available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    if "torch._six" not in sys.modules:
        shim = types.ModuleType("torch._six")
        shim.builtins = builtins
        sys.modules["torch._six"] = shim
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "utils" or k.startswith("utils.")}
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import utils.core as ref_core          # noqa: E402
        import utils.attacks as ref_attacks    # noqa: E402
    finally:
        sys.path.remove(REFERENCE_ROOT)
        # keep the reference importable under a private alias, restore whatever was there
        for k in list(sys.modules):
            if k == "utils" or k.startswith("utils."):
                sys.modules["_ee_reference_" + k] = sys.modules.pop(k)
        sys.modules.update(saved)
    return ref_core, ref_attacks


@contextlib.contextmanager
def quiet():
    """The reference constructors print a banner (core.py:160); keep test logs clean."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield
