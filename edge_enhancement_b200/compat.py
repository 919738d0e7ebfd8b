"""Script-level stand-ins so that the reference's experiment scripts import and run on a box
that lacks their small third-party dependencies.  Registered by ``edge_enhancement_b200.install(shims=True)`` ONLY for
modules that are not importable; nothing here is on the hot path.

* ``easydict.EasyDict``  -- attribute-style dict (utils/helper.py:9, :115-127 builds ``args`` with it)
* ``managpu.GpuManager`` -- the scripts call ``GpuManager().set_by_memory(1)`` at import time to pick a GPU
  (MNIST/experiments_mnist.py:20-22); here the current CUDA device is kept
* ``autoattack.AutoAttack`` -- imported at the top of the MNIST / Tiny-ImageNet scripts but only used by ``validate_aa``;
  the stand-in raises when it is actually run
* ``turtle`` -- a stray editor auto-import in one reference model file; needs tkinter, never used
"""
import sys
import types


class EasyDict(dict):
    """dict whose keys are also attributes; nested dicts (and dicts inside lists / tuples) are converted recursively."""

    def __init__(self, d=None, **kwargs):
        super().__init__()
        d = dict(d or {})
        d.update(kwargs)
        for k, v in d.items():
            setattr(self, k, v)

    @classmethod
    def _convert(cls, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._convert(x) for x in v)
        return v

    def __setattr__(self, name, value):
        value = self._convert(value)
        super().__setattr__(name, value)
        super().__setitem__(name, value)

    __setitem__ = __setattr__

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def update(self, e=None, **f):
        d = dict(e or {})
        d.update(f)
        for k, v in d.items():
            setattr(self, k, v)


class GpuManager:
    """stand-in for managpu.GpuManager: keeps the process on its current CUDA device"""

    def __init__(self, *args, **kwargs):
        pass

    def set_by_memory(self, n=1, *args, **kwargs):
        try:
            import torch
            return [torch.cuda.current_device()][:n] if torch.cuda.is_available() else []
        except Exception:
            return []


class AutoAttack:
    """stand-in for autoattack.AutoAttack: constructing it is allowed (the scripts do so lazily), running it is not"""

    def __init__(self, *args, **kwargs):
        self.args, self.kwargs = args, kwargs

    def run_standard_evaluation(self, *args, **kwargs):
        raise RuntimeError("autoattack is not installed: AutoAttack evaluation (validate_aa) is outside edge_enhancement_b200")


def install_shims():
    """Register the stand-ins for every module of (easydict, managpu, autoattack) that cannot be imported.
    Returns the list of module names that were shimmed."""
    import importlib
    done = []
    # `turtle`: Tiny_ImageNet/models_tinyimagenet/resnet_EE_square.py:5 carries a stray `from turtle import forward`
    # (an editor auto-import); it fails on headless boxes without tkinter and the name is never used
    for name, attrs in (("easydict", {"EasyDict": EasyDict}), ("managpu", {"GpuManager": GpuManager}),
                        ("autoattack", {"AutoAttack": AutoAttack}), ("turtle", {"forward": lambda *a, **k: None})):
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
        except ImportError:
            m = types.ModuleType(name)
            m.__dict__.update(attrs)
            m.__edge_b200_shim__ = True
            sys.modules[name] = m
            done.append(name)
    return done
