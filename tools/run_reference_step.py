#!/usr/bin/env python
"""Run the reference's OWN experiment scripts (their unmodified main(): model construction from the reference's model
files, train() over a synthetic loader, validate(), save_checkpoint) on top of this repo's drop-ins.

    EE_REFERENCE_ROOT=/path/to/Edge-Enhancement python tools/run_reference_step.py mnist    [--config ee_at_training.yml]
    EE_REFERENCE_ROOT=...                         python tools/run_reference_step.py tiny     [--config ee_at_bpda3_square.yml]
    EE_REFERENCE_ROOT=...                         python tools/run_reference_step.py imagenet [--config at_ee_training.yml]
    ... --dry-run     # CPU-only box: everything up to the first kernel call, which must be the drop-in's "no CPU fallback"

SURVEY.md section 8(f)-4.  `edge_enhancement_b200.install(shims=True)` makes `from utils.core import ...` /
`from utils.attacks import ...` in the scripts AND in the reference's model files resolve to the drop-ins, and provides
stand-ins for easydict / managpu / autoattack.  Nothing in the reference tree is edited; the things that keep the scripts
from running as shipped are patched on the imported module objects, each patch listed in PATCHES:

  * data loaders      -> a few synthetic batches of the right shape (no dataset on the box);
  * config defaults   -> keys the scripts read but the YAMLs lack (`type_canny` for MNIST, `step_size_3` shadowed by the
                         duplicate `step_size_1` key, `n_queries`, `beta`, `cize`, `nGPU`, ...); duplicate YAML keys are reported;
  * one epoch, workers 0, print every batch;
  * ImageNet          -> `validate` is called with 6 arguments but defined with 8 (experiments_imagenet.py:185,196 vs :300):
                         the module-level name is wrapped so that the missing num_steps / step_size come from the config;
                         a single-process `env://` process group is set up for its DistributedDataParallel.

The reference checkout does not travel to the GPU box of this project's driver, so a recorded run needs a box where both are
present; tests/test_host_logic.py runs the --dry-run mode in the build container.
"""
import argparse
import importlib
import importlib.util
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SCRIPTS = {
    # name: (directory, script module, config dir, default config, loader name, image shape, classes)
    "mnist": ("MNIST", "experiments_mnist", "configs_mnist", "ee_at_training.yml", "data_loader_mnist", (1, 28, 28), 10),
    "tiny": ("Tiny_ImageNet", "experiments_tinyimagenet", "configs_tinyimagenet", "ee_at_training.yml", "data_loader_tiny_imagenet", (3, 64, 64), 200),
    "imagenet": ("ImageNet", "experiments_imagenet", "configs_imagenet", "at_ee_training.yml", "data_loader_imagenet_dataset", (3, 224, 224), 1000),
}

DEFAULTS = {   # read by the scripts, absent from (some of) the YAMLs
    "type_canny": "CannyFilter", "n_queries": 1, "beta": 6.0, "cize": None, "nGPU": 1, "local_rank": 0, "gf": False,
    "alpha": 0.0, "sigma": 1.0, "pretrained": False, "resume": "", "evaluate": False, "attack_method": "PGD", "no_cuda": False,
    "prob_start_from_clean": 0.0, "label_smoothing": 0.0,
}

PATCHES = []


def note(msg):
    PATCHES.append(msg)
    print("[run_reference_step] " + msg, flush=True)


def duplicate_yaml_keys(path):
    """Top-level keys that appear more than once (PyYAML keeps the last one silently, e.g. `step_size_1` twice in
    Tiny_ImageNet/configs_tinyimagenet/ee_at_training.yml:29,37, which leaves `step_size_3` undefined)."""
    seen, dup = set(), []
    with open(path) as f:
        for line in f:
            if line[:1] in (" ", "\t", "#", "\n", "-"):
                continue
            key = line.split(":", 1)[0].strip()
            if key in seen:
                dup.append(key)
            seen.add(key)
    return dup


def synthetic_loader(shape, n_class, batch, batches):
    import torch
    g = torch.Generator().manual_seed(0)
    x = torch.rand((batch * batches,) + tuple(shape), generator=g)
    y = torch.randint(0, n_class, (batch * batches,), generator=g)
    return torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=batch, shuffle=False, num_workers=0)


def _cpu_torch_proxy(torch):
    """For the ImageNet script's dry run on a box without GPUs: main() hard-codes torch.device("cuda", rank), the NCCL
    backend and DistributedDataParallel(device_ids=[rank]).  A proxy of the `torch` module, installed as the SCRIPT's global
    name only, maps those three to their CPU equivalents so that main() still runs up to the first kernel call."""
    class Proxy(object):
        def __init__(self, target, overrides):
            self._t, self._o = target, overrides

        def __getattr__(self, n):
            return self._o[n] if n in self._o else getattr(self._t, n)

    def init_pg(backend=None, **kw):
        return torch.distributed.init_process_group(backend="gloo", **kw)

    def ddp(model, device_ids=None, output_device=None, **kw):
        return torch.nn.parallel.DistributedDataParallel(model, **kw)
    dist = Proxy(torch.distributed, {"init_process_group": init_pg})
    parallel = Proxy(torch.nn.parallel, {"DistributedDataParallel": ddp})
    sync_bn = Proxy(torch.nn.SyncBatchNorm, {"convert_sync_batchnorm": lambda m: m})        # SyncBatchNorm needs GPU modules
    nn = Proxy(torch.nn, {"parallel": parallel, "SyncBatchNorm": sync_bn})
    cuda = Proxy(torch.cuda, {"set_device": lambda *a, **k: None})
    return Proxy(torch, {"device": lambda *a, **k: torch.device("cpu"), "distributed": dist, "nn": nn, "cuda": cuda})


def prepare(name, config, batch, batches, dry_run):
    """Import the script on top of the drop-ins and patch the module object.  Returns (module, config path)."""
    ref = os.environ.get("EE_REFERENCE_ROOT", "/root/reference")
    d, modname, cfgdir, default_cfg, loader_name, shape, n_class = SCRIPTS[name]
    script_dir = os.path.join(ref, d)
    if not os.path.isfile(os.path.join(script_dir, modname + ".py")):
        raise SystemExit("reference script not found under %s (set EE_REFERENCE_ROOT)" % script_dir)
    cfg_path = os.path.join(script_dir, cfgdir, config or default_cfg)
    for p in (ref, script_dir):
        if p not in sys.path:
            sys.path.insert(0, p)
    for k in [k for k in sys.modules if k == "utils" or k.startswith("utils.")]:
        del sys.modules[k]

    import torch
    if "torch._six" not in sys.modules:            # the reference's vendored utils/_jit_internal.py:9
        import builtins
        import types
        shim = types.ModuleType("torch._six")
        shim.builtins = builtins
        sys.modules["torch._six"] = shim
    import edge_enhancement_b200 as ee
    ee.install(shims=True)
    if dry_run and not torch.cuda.is_available():
        # the reference's model files call .cuda() in constructors (u2net.Sobel, core.Add_Square): identity on a CPU box
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
        note("dry run on a CPU box: Tensor.cuda / Module.cuda are the identity")

    mod = importlib.import_module(modname)
    dup = duplicate_yaml_keys(cfg_path)
    if dup:
        note("config %s has duplicate keys %s (PyYAML keeps the last value, like in the reference's own runs)" % (os.path.basename(cfg_path), dup))

    orig_parse = mod.parse_config_file

    def parse_config_file(args):
        cfg = orig_parse(args)
        for k, v in DEFAULTS.items():
            if k not in cfg:
                cfg[k] = v
                note("config default %s = %r" % (k, v))
        if cfg.get("cize") is None:
            cfg["cize"] = shape[-1]
        for i in (2, 3):                              # the PGD-50 / PGD-100 evaluation settings shadowed by duplicate keys
            cfg.setdefault("num_steps_%d" % i, cfg["num_steps_1"])
            cfg.setdefault("step_size_%d" % i, cfg["step_size_1"])
        cfg["epochs"] = cfg.get("start_epoch", 0) + 1
        cfg["workers"] = 0
        cfg["print_freq"] = 1
        cfg["batch_size"] = batch * (cfg.get("nGPU", 1) if name == "imagenet" else 1)
        return cfg
    mod.parse_config_file = parse_config_file
    note("one epoch, workers 0, batch %d, %d synthetic batches" % (batch, batches))

    loader = synthetic_loader(shape, n_class, batch, batches)
    if name == "imagenet":
        setattr(mod, loader_name, lambda data: (loader.dataset, loader.dataset))
        orig_validate = mod.validate

        def validate(val_loader, model, criterion, print_freq, device, *rest):
            if len(rest) == 1:                      # experiments_imagenet.py:185,196 pass 6 arguments, :300 takes 8
                return orig_validate(val_loader, model, criterion, print_freq, device, mod.args.num_steps_1, mod.args.step_size_1, rest[0])
            return orig_validate(val_loader, model, criterion, print_freq, device, *rest)
        mod.validate = validate
        note("validate(): 6-argument calls get num_steps_1 / step_size_1 from the config")
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        os.environ.setdefault("RANK", "0")
        os.environ.setdefault("WORLD_SIZE", "1")
        if dry_run and not torch.cuda.is_available():
            mod.torch = _cpu_torch_proxy(torch)
            mod.dist = mod.torch.distributed
            note("dry run on a CPU box: the script's `torch` name is a proxy (device -> cpu, process group -> gloo, DDP without device ids)")
    else:
        setattr(mod, loader_name, lambda *a, **k: (loader, loader))
    note("%s -> synthetic loader of %s images, %d classes" % (loader_name, "x".join(map(str, shape)), n_class))
    return mod, cfg_path


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("script", choices=sorted(SCRIPTS))
    ap.add_argument("--config", default=None, help="file name inside the script's config directory")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--batches", type=int, default=2)
    ap.add_argument("--dry-run", action="store_true", help="CPU box: succeed when main() reaches the drop-in's 'no CPU fallback' error")
    a = ap.parse_args()
    mod, cfg_path = prepare(a.script, a.config, a.batch, a.batches, a.dry_run)
    argv = [mod.__file__, "--config", cfg_path, "--data", "synthetic"]
    if a.script == "imagenet":
        argv += ["--local_rank", "0"]
    out_dir = tempfile.mkdtemp(prefix="ee_ref_step_")
    cwd = os.getcwd()
    os.chdir(out_dir)
    sys.argv = argv
    t0 = time.time()
    try:
        mod.main()
    except RuntimeError as e:
        if a.dry_run and "no CPU fallback" in str(e):
            print("[run_reference_step] DRY RUN OK: %s.main() built its model from the reference's files and reached the drop-in "
                  "hot path: %s" % (SCRIPTS[a.script][1], e))
            return 0
        raise
    finally:
        os.chdir(cwd)
    print("[run_reference_step] OK: %s.main() ran one epoch of train() + validate() on the drop-ins in %.1f s (outputs under %s); "
          "%d patches applied" % (SCRIPTS[a.script][1], time.time() - t0, out_dir, len(PATCHES)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
