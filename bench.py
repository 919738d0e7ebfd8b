#!/usr/bin/env python
"""bench.py -- edge-enhanced PGD-10 hot-path throughput (images/s) on N B200s, with the HBM roofline
of the dominant kernel and the CPU path timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W]          # our CUDA path (default N=1)
    python bench.py --impl reference ...                         # the CPU port of the reference path
    torchrun --nproc-per-node N bench.py --gpus N ...            # one rank per GPU, batch sharded

Workload (BASELINE.json configs[1], SURVEY.md section 8d "C2"): Tiny-ImageNet edge-enhanced PGD-10,
3x64x64 fp32, CannyFilter_step125_1 (high 76/255, w=1), eps 16/255, step 2/255
(Tiny_ImageNet/configs_tinyimagenet/ee_at_bpda3_square.yml).  One STEP is the hot path of one
PGD-10 adversarial-training iteration over the resident batch:
    10 x [ edge+blend forward -> edge+blend backward -> PGD L-inf step ]  +  1 final forward
i.e. 31 launches of our kernels.  The CNN forward/backward between them (cuDNN) and the FFT low-pass
(cuFFT) are outside the product: `base` (x_hfs) and `g_out` (dL/d blended image) are resident
synthetic tensors.  Per rank the step processes `--batch` images (default 16 Tiny batches of 256 =
4096 images, 201 MB per tensor, so every launch streams from HBM rather than the 126 MB L2).

Timing: W warm-up steps, then exactly K steps between barrier+synchronize, CUDA events on the
launching stream, max over ranks.  `value` = images of all ranks / that time with inputs resident
in HBM.  `e2e` = the same 31-launch step issued through the public functional API (= the C-ABI entry
points) with pinned HOST buffers: H2D of the clean batch and a device->host read of the step's result (a
per-image metric of the blended adversarial image) inside the timed region; `e2e.full_readback` also
returns the whole adversarial batch to the host (round 1's definition).  Both are reported next to the
pure-CUDA copy floor of the same traffic.  `e2e_attack_api` = attacks.PGD (the reference's call
signature) on an edge_enhance front end + a torch stand-in head, same host buffers; the stand-in's torch
kernels and autograd are inside that number.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# algorithmic bytes per pixel (C = 3, fp32), SURVEY.md section 8d / BASELINE.md section 2
BYTES_FWD_PX = 36.0      # read x, base ; write out
BYTES_BWD_PX = 60.0      # read g_out, x, base ; write g_x, g_base
BYTES_PGD_ELT = 16.0     # read x, g, x0 ; write x'

EPS, ALPHA, HIGH, LOW, W_BLEND = 16 / 255, 2 / 255, 76 / 255, 38 / 255, 1.0
N_PGD = 10


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="images per rank per step")
    ap.add_argument("--side", type=int, default=64)
    ap.add_argument("--variant", default="step125", choices=["step125", "canny", "bpda"])
    ap.add_argument("--cpu-images", type=int, default=0, help="images per step of the CPU arm (0 = --batch, the GPU arm's step)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work in the cpu_baseline leg (bounded sample)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=4, help="chunks the e2e batch is pipelined in")
    ap.add_argument("--e2e-streams", type=int, default=4, help="CUDA streams the e2e chunks are issued on round-robin")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--th-fwd", type=int, default=0)
    ap.add_argument("--th-bwd", type=int, default=0)
    ap.add_argument("--sweep", action="store_true", help="extra: kernel sweep table on stderr (configs[4])")
    ap.add_argument("--no-named-batch", action="store_true", help="skip the named_batch block (per-launch times at the configs' own batch sizes)")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra blocks (configs[0], [2], [3] with real CNNs)")
    args = ap.parse_args()
    if args.cpu_images <= 0:
        args.cpu_images = args.batch
    return args


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def bind_to_gpu_numa_node(index):
    """Pin this rank's host threads (and therefore the first-touch placement of its pinned staging buffers) to the CPUs
    NVML reports as local to GPU `index`, so that with several ranks per box the H2D / D2H traffic of the e2e leg does
    not cross the socket interconnect.  Returns the number of CPUs bound to (0 = left unchanged)."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(index)
        words = nv.nvmlDeviceGetCpuAffinity(h, ((os.cpu_count() or 64) + 63) // 64)
        cpus = [i * 64 + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def shard_sizes(total, world):
    """Contiguous batch shards, sizes differ by at most one (what DistributedSampler does)."""
    base, rem = divmod(total, world)
    return [base + (1 if r < rem else 0) for r in range(world)]


def ncu_traffic(kernel, shape, variant):
    """DRAM bytes per launch of `kernel` from the committed ncu capture, if it was taken on this shape."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)
        if list(t["shape"]) == list(shape) and t["variant"] == variant:
            return t[kernel]["dram_read_bytes"] + t[kernel]["dram_write_bytes"]
    except Exception:
        pass
    return None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# clocks sampled DURING the timed region
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons with a separate `nvidia-smi -lms` process (a Python thread is
    starved by the launch loop) and keeps the samples whose timestamps fall inside the timed region."""
    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_ms=20):
        import subprocess
        import tempfile
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        self.t0 = self.t1 = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", str(period_ms)],
                                         stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def begin(self):
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()

    def stop(self):
        import datetime
        time.sleep(0.05)
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        rows = []
        try:
            self.tmp.flush()
            with open(self.tmp.name) as f:
                for line in f:
                    parts = [p.strip() for p in line.split(",")]
                    if len(parts) < 7:
                        continue
                    try:
                        ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                        rows.append((ts, float(parts[1]), float(parts[2]), parts[3:7]))
                    except ValueError:
                        continue
            os.unlink(self.tmp.name)
        except OSError:
            pass
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": "nvidia-smi unavailable"}
        inside = [r for r in rows if self.t0 is not None and self.t0 <= r[0] <= self.t1]
        note = "inside the timed region"
        if not inside:          # region shorter than the sampling period: take the samples closest to it
            mid = 0.5 * ((self.t0 or rows[-1][0]) + (self.t1 or rows[-1][0]))
            inside = sorted(rows, key=lambda r: abs(r[0] - mid))[:3]
            note = "nearest samples (timed region shorter than the sampling period)"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in inside for n, v in zip(names, r[3]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median([r[1] for r in inside])), "sm_max_mhz": inside[0][2], "reasons": reasons,
                "samples": len(inside), "source": "nvidia-smi -lms 20, " + note}


def sample_clocks_until(index, done_event, period=0.002):
    """NVML clock / throttle-reason samples taken by the launching thread while the already enqueued
    timed steps are still executing (launches run far ahead of the GPU), i.e. under load."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(index)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
    except Exception as e:
        return {"samples": 0, "error": "nvml init: %r" % (e,)}
    names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
             0x80: "hw_power_brake"}
    samples, reasons, err = [], set(), None
    t_stop = time.time() + 120.0
    while time.time() < t_stop:
        try:
            finished = bool(done_event.query())
        except Exception as e:          # pragma: no cover
            err = "event.query: %r" % (e,)
            break
        try:
            samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
            try:
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            except Exception:
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            reasons |= {n for bit, n in names.items() if mask & bit}
        except Exception as e:
            err = "nvml query: %r" % (e,)
            break
        if finished:                    # the last sample may fall just after the end; it is still recorded
            break
        time.sleep(period)
    out = {"sm_mhz": float(np.median(samples)) if samples else None, "sm_max_mhz": float(mx), "reasons": sorted(reasons),
           "samples": len(samples), "source": "NVML polled by the launching thread while the timed steps execute"}
    if err:
        out["error"] = err
    return out


# ---------------------------------------------------------------------------------------------
# CPU path (oracle port of the reference): the same PGD-10 hot path on a bounded sample
# ---------------------------------------------------------------------------------------------
def cpu_hot_path(images, side, variant, repeats=1, min_seconds=0.0):
    """Returns (images/s, cores, seconds per step) of the oracle port running the PGD-10 hot-path step on `images`
    images: `repeats` steps, or as many as fit in `min_seconds` of CPU work (whichever is more); mean over the steps."""
    from oracle import oracle as O
    r = np.random.default_rng(1234)
    shape = (images, 3, side, side)
    x0 = r.random(shape, dtype=np.float32)
    base = (r.random(shape, dtype=np.float32) * 1.1 - 0.1).astype(np.float32)
    g_out = r.standard_normal(shape, dtype=np.float32)
    p = O.make_params(variant, alpha=0.0, low=None if variant == "step125" else LOW, high=HIGH, hysteresis=True)
    cores = O.threads()
    x_start = np.clip(x0 + (r.random(shape, dtype=np.float32) * 2 - 1) * np.float32(EPS), 0, 1).astype(np.float32)
    total, n = 0.0, 0
    while n < repeats or total < min_seconds:
        x = x_start
        t0 = time.perf_counter()
        for _it in range(N_PGD):
            O.edge_blend_fwd(x, base, p, W_BLEND)
            g_x, _g_base = O.edge_blend_bwd(g_out, x, base, p, W_BLEND)
            x = O.pgd_linf_step(x, g_x, x0, ALPHA, EPS)
        O.edge_blend_fwd(x, base, p, W_BLEND)
        total += time.perf_counter() - t0
        n += 1
    cpu_hot_path.last_steps = n
    return images * n / total, cores, total / n


def cpu_live_reference(images, side, variant, seconds=8.0):
    """The SAME PGD-10 hot-path step through the reference's own torch modules on the host CPU (kind "live") -- only where
    the reference checkout is present (the build container; it cannot travel to the GPU box).  Returns None otherwise."""
    try:
        from oracle import ref_loader
        if not ref_loader.available():
            return None
        import torch
        rc, _ra = ref_loader.load()
    except Exception:
        return None
    torch.set_num_threads(os.cpu_count() or 1)
    cls = {"step125": "CannyFilter_step125_1", "canny": "CannyFilter", "bpda": "CannyFilter_BPDA"}[variant]
    with ref_loader.quiet():
        f = getattr(rc, cls)(use_cuda=False, alpha=0.0)
    g = torch.Generator().manual_seed(1234)
    shape = (images, 3, side, side)
    x0 = torch.rand(shape, generator=g)
    base = torch.rand(shape, generator=g) * 1.1 - 0.1
    g_out = torch.randn(shape, generator=g)
    low = None if variant == "step125" else LOW

    def step():
        x = x0.clone()
        for _ in range(N_PGD):
            x.requires_grad_()
            out = torch.clamp(base + W_BLEND * f(x, low_threshold=low, high_threshold=HIGH, hysteresis=True), 0.0, 1.0)
            grad = torch.autograd.grad(out, [x], grad_outputs=g_out)[0]
            grad = torch.nan_to_num(grad)
            x = x.detach() + ALPHA * torch.sign(grad)                              # utils/attacks.py:25-27
            x = torch.min(torch.max(x, x0 - EPS), x0 + EPS)
            x = torch.clamp(x, 0, 1)
        with torch.no_grad():
            torch.clamp(base + W_BLEND * f(x, low_threshold=low, high_threshold=HIGH, hysteresis=True), 0.0, 1.0)

    step()
    total, n = 0.0, 0
    while total < seconds:
        t0 = time.perf_counter()
        step()
        total += time.perf_counter() - t0
        n += 1
    return {"value": images * n / total, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "live",
            "sample": "%d images of 3x%dx%d per step, %d PGD-10 hot-path steps through the reference's own torch modules" % (images, side, side, n)}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference is Python
    (torch eager) and cannot travel to the GPU box, so this is the C port under oracle/ (OpenMP, all
    host threads), each step a bounded sample of the same workload."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm is the CPU path on ALL host cores
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    images = args.cpu_images
    for _ in range(max(args.warmup, 0)):
        cpu_hot_path(images, args.side, args.variant)
    dt = 0.0
    for _ in range(args.steps):
        dt += cpu_hot_path(images, args.side, args.variant)[2]     # hot path only, not the input synthesis
    value = images * args.steps / dt
    from oracle import oracle as O
    cores = O.threads()
    line = {
        "impl": "reference", "metric": "edge-enhanced PGD-10 hot-path images/sec", "value": value, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, images, 1, cpu=True),
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": "%d images of 3x%dx%d per step (the GPU arm's step), full PGD-10 hot path"
                                   % (images, args.side, args.side)},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, per_rank, world, cpu=False):
    return {
        "workload": "configs[1]: Tiny-ImageNet edge-enhanced PGD-10 hot path, 3x%dx%d fp32, %s, eps 16/255, step 2/255; "
                    "one step = 10x(edge+blend fwd, edge+blend bwd, PGD L-inf step) + final fwd = 31 launches; "
                    "CNN and FFT low-pass are outside the product (base, g_out resident synthetic tensors)"
                    % (args.side, args.side, {"step125": "CannyFilter_step125_1", "canny": "CannyFilter", "bpda": "CannyFilter_BPDA"}[args.variant]),
        "images_per_rank_per_step": per_rank,
        "global_images_per_step": per_rank * world,
        "tiny_batches_of_256_per_rank": per_rank / 256.0,
        "l2": ("n/a (CPU)" if cpu else
               "inputs larger than L2: %.0f MB per tensor, 7 tensors per rank, vs 126 MB L2" % (per_rank * 3 * args.side * args.side * 4 / 1e6)),
        "parallelism": "batch sharded over %d rank(s), no data-path collective" % world,
    }



# ---------------------------------------------------------------------------------------------
# named_batch: the hot-path kernels at the batch sizes the configs really name (launch-bound there)
# ---------------------------------------------------------------------------------------------
NAMED_SHAPES = (
    # tag, B, C, side, variant, alpha, low, high          (sizes M, T, I)
    ("configs[1] T 256x3x64x64 step125", 256, 3, 64, "step125", 0.0, None, 76 / 255),
    ("configs[1] T 256x3x64x64 canny", 256, 3, 64, "canny", 0.0, 38 / 255, 76 / 255),
    ("configs[0] M 128x1x28x28 canny a=0.3", 128, 1, 28, "canny", 0.3, 25 / 255, 51 / 255),
    ("configs[3] I 32x3x224x224 step125", 32, 3, 224, "step125", 0.0, None, 76 / 255),
)


def _graph_us_per_launch(torch, fn, reps=20, replays=10):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / (replays * reps)


def run_named_batch(torch, F_ee, core, dev, peak):
    """us per launch of fwd / bwd / PGD step / the single-call iteration at the configs' OWN batch sizes: through the
    eager API (Python + ctypes + launch; host-bound at these sizes) and as a CUDA-graph replay of 20 launches (what
    attacks.GraphedPGD does).  These tensors are L2-resident, so GB/s may exceed the HBM peak; frac is reported against the
    same HBM peak as everything else."""
    import contextlib, io
    rows = []
    for tag, B, C, S, variant, alpha, low, high in NAMED_SHAPES:
        cls = {"step125": core.CannyFilter_step125_1, "canny": core.CannyFilter, "bpda": core.CannyFilter_BPDA}[variant]
        with contextlib.redirect_stdout(io.StringIO()):
            p = cls(use_cuda=False, alpha=alpha).params(low, high, True)
        gen = torch.Generator(device=dev).manual_seed(99)
        shape = (B, C, S, S)
        x = torch.rand(shape, device=dev, generator=gen); base = torch.rand(shape, device=dev, generator=gen) * 1.1 - 0.1
        g = torch.randn(shape, device=dev, generator=gen); x0 = torch.rand(shape, device=dev, generator=gen)
        o1, o2, o3, o4 = (torch.empty_like(x) for _ in range(4))
        npx = B * S * S
        ops = (("edge_blend_fwd", lambda: F_ee.edge_blend(x, base, p, 1.0, out=o1), 12.0 * C * npx),
               ("edge_blend_bwd", lambda: F_ee.edge_blend_backward(g, x, base, p, 1.0, g_x=o2, g_base=o3), 20.0 * C * npx),
               ("pgd_linf_step", lambda: F_ee.pgd_linf_step(x, g, x0, ALPHA, EPS, out=o4), 16.0 * C * npx),
               ("iteration_one_call", lambda: F_ee.pgd_iteration(x, base, g, x0, p, 1.0, ALPHA, EPS, out=o1, g_x=o2, g_base=o3, x_next=o4),
                48.0 * C * npx))
        row = {"shape": tag, "mb_per_tensor": 4.0 * C * npx / 1e6}
        for name, fn, nbytes in ops:
            for _ in range(5):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 200
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            eager_us = 1e3 * e0.elapsed_time(e1) / n
            graph_us = _graph_us_per_launch(torch, fn)
            row[name] = {"eager_us": round(eager_us, 2), "graph_us": round(graph_us, 2),
                         "graph_gbs": round(nbytes / graph_us / 1e3, 1), "graph_frac_of_hbm_peak": round(nbytes / graph_us / 1e3 / peak, 3)}
        rows.append(row)
    return {"note": "us per launch at the configs' own batch sizes; eager = public API call in a Python loop (host-bound), "
                    "graph = CUDA-graph replay of 20 such launches; tensors are L2-resident (<= 19 MB)", "rows": rows}


# ---------------------------------------------------------------------------------------------
# extra: configs[0], [2], [3] with real (stock torch) CNNs behind the drop-in front end
# ---------------------------------------------------------------------------------------------
def run_extra_configs(torch, core, attacks, dev):
    """Driver-visible numbers for the other BASELINE configs (the judged metric stays configs[1]).  The CNNs are stock torch
    (outside the product); the front end (HighFreqSuppress + edge filter + blend = core.EdgeEnhance) and the attack updates
    are this repo's kernels.  Synthetic data, random-init weights, one GPU."""
    import contextlib, io
    import torch.nn as nn
    import torch.nn.functional as Fn
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import train_throughput as TT
    out = {}

    def timed(fn, n):
        fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / n

    def front(size, r, w, low, high, alpha, kind):
        with contextlib.redirect_stdout(io.StringIO()):
            return core.EdgeEnhance(cize=size, r=r, w=w, low=low, high=high, alpha=alpha, type_canny=kind).to(dev)

    # ---- configs[0]: MNIST small CNN, edge-enhanced PGD-40 eval, batch 128 (MNIST/experiments_mnist.py:271-304; Net2_EE)
    try:
        class Net2(nn.Module):
            def __init__(self, fr):
                super().__init__()
                self.front = fr
                self.c1, self.c2 = nn.Conv2d(1, 32, 5, padding=2), nn.Conv2d(32, 64, 5, padding=2)
                self.f1, self.f2 = nn.Linear(64 * 7 * 7, 1024), nn.Linear(1024, 10)

            def forward(self, x):
                x = self.front(x)
                x = Fn.max_pool2d(Fn.relu(self.c1(x)), 2)
                x = Fn.max_pool2d(Fn.relu(self.c2(x)), 2)
                return self.f2(Fn.relu(self.f1(x.flatten(1))))

        class A0:
            random, epsilon = True, 0.3
        m = Net2(front(28, 4, 1.0, 25.0, 51.0, 0.3, 'CannyFilter')).to(dev).eval()
        x = torch.rand(128, 1, 28, 28, device=dev); y = torch.randint(0, 10, (128,), device=dev)
        ms_eager = timed(lambda: attacks.PGD(m, A0, x, y, 40, 0.01), 3)
        gp = attacks.GraphedPGD(m, A0, x, y, 0.01)
        ms_graph = timed(lambda: gp(x, y, 40), 5)
        same = bool(torch.equal(_seeded(torch, lambda: attacks.PGD(m, A0, x, y, 40, 0.01)), _seeded(torch, lambda: gp(x, y, 40))))
        out["configs[0]"] = {"workload": "MNIST Net2-style CNN + EdgeEnhance(28, r 4, CannyFilter alpha 0.3, 25/51, w 1), PGD-40 eval attack, "
                                         "batch 128, eps 0.3, step 0.01, random start", "attack_ms_eager": ms_eager, "attack_ms_graphed": ms_graph,
                             "images_per_s_eager": 128 / ms_eager * 1e3, "images_per_s_graphed": 128 / ms_graph * 1e3,
                             "graphed_equals_eager_bitwise": same}
        del m, gp
    except Exception as e:      # pragma: no cover
        out["configs[0]"] = {"error": repr(e)}

    # ---- configs[2]: Tiny-ImageNet TRADES, KL inner loop, batch 256 (utils/attacks.py:404-418)
    try:
        class A2:
            random, epsilon = False, 16 / 255
        m = TT.PreActResNet18(front(64, 8, 1.0, 38.0, 76.0, 0.0, 'CannyFilter_step125_1')).to(dev).eval()
        x = torch.rand(256, 3, 64, 64, device=dev)
        tr = attacks.Trades(step_size=2 / 255, epsilon=16 / 255, perturb_steps=10, beta=6.0)
        with torch.no_grad():
            logits = m(x)
        ms_eager = timed(lambda: tr.PGD_Linf(m, x, logits), 2)
        kl = nn.KLDivLoss(reduction="batchmean")
        prob = Fn.softmax(logits, dim=-1)
        gp = attacks.GraphedPGD(m, A2, x, prob, 2 / 255, loss_fn=lambda lg, pr: kl(Fn.log_softmax(lg, dim=1), pr))
        ms_graph = timed(lambda: gp(x, prob, 10, x_init=x + 0.001 * torch.randn_like(x)), 3)
        out["configs[2]"] = {"workload": "Tiny-ImageNet PreAct-ResNet18 + EdgeEnhance(64, r 8, step125, 76, w 1), TRADES KL inner loop "
                                         "(Trades.PGD_Linf, 10 steps, eps 16/255, step 2/255), batch 256", "attack_ms_eager": ms_eager,
                             "attack_ms_graphed": ms_graph, "images_per_s_eager": 256 / ms_eager * 1e3, "images_per_s_graphed": 256 / ms_graph * 1e3}
        del m, gp
    except Exception as e:      # pragma: no cover
        out["configs[2]"] = {"error": repr(e)}

    # ---- configs[3]: ImageNet ResNet-50 free adversarial training, 32 images per GPU (AT_hfs_canny_free_imagenet_ddp.py:311-334)
    try:
        m = TT.ResNet50(front(224, 16, 1.0, 38.0, 76.0, 0.0, 'CannyFilter_step125_1')).to(dev).train()
        opt = torch.optim.SGD(m.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
        x = torch.rand(32, 3, 224, 224, device=dev); y = torch.randint(0, 1000, (32,), device=dev)
        noise = torch.zeros(32, 3, 224, 224, device=dev)
        state = {"inp": x.clone()}

        def free_iteration():
            for _ in range(4):                                     # n_repeats = 4
                inp = state["inp"].detach().requires_grad_()
                loss = Fn.cross_entropy(m(inp), y)
                opt.zero_grad(set_to_none=True)
                loss.backward()
                state["inp"] = attacks.free_at_update_(noise, inp.grad, x, 4 / 255, 4 / 255)
                opt.step()
        ms = timed(free_iteration, 3)
        out["configs[3]"] = {"workload": "ImageNet ResNet-50 + EdgeEnhance(224, r 16, step125, 76, w 1), free adversarial training "
                                         "(n_repeats 4, clip_eps = fgsm_step = 4/255), 32 images per GPU, SGD; one GPU, no DDP",
                             "ms_per_minibatch_4_repeats": ms, "images_per_s": 32 / ms * 1e3, "gradient_steps_per_s": 4 / ms * 1e3}
        del m, opt
    except Exception as e:      # pragma: no cover
        out["configs[3]"] = {"error": repr(e)}
    torch.cuda.empty_cache()
    # ---- the low-pass of the *_EE front end at the bench batch: exact FFMA kernel (default) and the tensor-core variant
    try:
        from edge_enhancement_b200 import functional as F_ee
        x = torch.rand((4096, 3, 64, 64), device=dev)
        y = torch.empty_like(x)
        hf = {}
        for impl in ("native", "tcgen05"):
            ms = timed(lambda: F_ee.hfs(x, 8, out=y, impl=impl), 10)
            hf[impl + "_us"] = ms * 1e3
            hf[impl + "_gbs"] = 8.0 * x.numel() / ms / 1e6
        exact = F_ee.hfs(x[:64], 8)
        hf["tcgen05_max_abs_diff_from_native"] = float((F_ee.hfs(x[:64], 8, impl="tcgen05") - exact).abs().max())
        hf["workload"] = "HighFreqSuppress(64, 64, 8) on 4096x3x64x64 fp32, 8 B/element algorithmic (ee_hfs_f32 / ee_hfs_tc_f32)"
        # the whole *_EE front end (low-pass + edge filter + blend) as one autograd node, forward + backward, on either low-pass
        g = torch.randn_like(x)
        for impl in ("native", "tcgen05"):
            with contextlib.redirect_stdout(io.StringIO()):
                fe = core.EdgeEnhance(cize=64, r=8, w=1.0, low=38.0, high=76.0, type_canny="CannyFilter_step125_1", hfs_impl=impl).to(dev)
            xr = x.clone().requires_grad_()

            def fwd_bwd():
                fe(xr).backward(g)
                xr.grad = None
            hf["front_end_fwd_bwd_us_" + impl] = timed(fwd_bwd, 5) * 1e3
            del xr
        out["highfreqsuppress"] = hf
        del x, y, g
    except Exception as e:      # pragma: no cover
        out["highfreqsuppress"] = {"error": repr(e)}
    torch.cuda.empty_cache()
    return out


def _seeded(torch, fn, seed=7):
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)
    return fn()


# ---------------------------------------------------------------------------------------------
# pure-CUDA copy floor of the e2e step (tools/copyfloor.cu)
# ---------------------------------------------------------------------------------------------
def copy_floor(torch, dist, dev, world, barrier, nbytes):
    """What this box delivers for the e2e step's traffic alone (one batch H2D + one batch D2H per step) with every rank
    copying at the same time: plain cudaMemcpyAsync from NUMA-local pinned buffers, no kernels, no torch.  max over ranks."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import copyfloor
        res = {}
        for key, mode, chunks, reg in (("h2d_only", 0, 1, False), ("d2h_only", 1, 1, False), ("both_1_chunk", 2, 1, False),
                                       ("both_4_chunks", 2, 4, False), ("both_1_chunk_host_register", 2, 1, True)):
            barrier()
            ms = copyfloor.measure(dev.index, nbytes, chunks=chunks, iters=10, mode=mode, host_register=reg)
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            res[key] = {"ms_per_step": ms, "gbs_each_way_per_rank": nbytes / ms / 1e6, "gbs_each_way_all_ranks": world * nbytes / ms / 1e6}
        torch.cuda.set_device(dev)
        return res
    except Exception as e:      # pragma: no cover
        return {"error": repr(e)}


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cpus = bind_to_gpu_numa_node(local) if world > 1 else 0
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=dev)

    import edge_enhancement_b200 as ee
    from edge_enhancement_b200 import functional as F_ee, _lib, core, attacks

    L = _lib.load()
    L.ee_set_tuning(args.th_fwd, args.th_bwd, 0)
    B, S = args.batch, args.side
    shape = (B, 3, S, S)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    x0 = torch.rand(shape, device=dev, generator=gen)
    base = torch.rand(shape, device=dev, generator=gen) * 1.1 - 0.1
    g_out = torch.randn(shape, device=dev, generator=gen)
    g_out.view(-1)[::97] = 0.0
    x_start = torch.clamp(x0 + (torch.rand(shape, device=dev, generator=gen) * 2 - 1) * EPS, 0, 1)
    xa, xb = torch.empty_like(x0), torch.empty_like(x0)
    out, g_x, g_base = torch.empty_like(x0), torch.empty_like(x0), torch.empty_like(x0)
    filt = {"step125": core.CannyFilter_step125_1, "canny": core.CannyFilter, "bpda": core.CannyFilter_BPDA}[args.variant]
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        canny = filt(use_cuda=False, alpha=0.0)
    p = canny.params(None if args.variant == "step125" else LOW, HIGH, True)

    stream = torch.cuda.current_stream(dev)
    ev_pool = []

    def step(record=None):
        """One PGD-10 hot-path step on resident tensors; 31 launches of our kernels."""
        cur = x_start
        nxt = xa
        for _it in range(N_PGD):
            if record is not None:
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e2 = torch.cuda.Event(enable_timing=True); e3 = torch.cuda.Event(enable_timing=True)
                e0.record(stream)
            F_ee.edge_blend(cur, base, p, W_BLEND, out=out)
            if record is not None:
                e1.record(stream)
            F_ee.edge_blend_backward(g_out, cur, base, p, W_BLEND, g_x=g_x, g_base=g_base)
            if record is not None:
                e2.record(stream)
            F_ee.pgd_linf_step(cur, g_x, x0, ALPHA, EPS, out=nxt)
            if record is not None:
                e3.record(stream)
                record.append((e0, e1, e2, e3))
            cur, nxt = nxt, (xb if nxt is xa else xa)
        F_ee.edge_blend(cur, base, p, W_BLEND, out=out)
        return 3 * N_PGD + 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    for _ in range(2):          # keep the GPU under load while nvidia-smi starts up
        step()
    records = []
    t_beg, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    barrier()
    sampler.begin()
    t_beg.record(stream)
    for _ in range(args.steps):
        launches += step(records)
    t_end.record(stream)
    nvml_samples = sample_clocks_until(local, t_end)      # the GPU is still draining the queued launches
    barrier()
    sampler.end()
    clocks = sampler.stop()
    if nvml_samples is not None and nvml_samples.get("samples", 0) >= max(2 if clocks.get("samples", 0) else 1, clocks.get("samples", 0)):
        clocks = nvml_samples
    elif nvml_samples is not None and nvml_samples.get("error"):
        clocks["nvml_error"] = nvml_samples["error"]
    ms = t_beg.elapsed_time(t_end)
    ms_t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_max = float(ms_t.item())
    value = world * B * args.steps / (ms_max / 1e3)

    # per-kernel durations inside the timed region
    fwd_ms = float(np.mean([a.elapsed_time(b) for a, b, _, _ in records]))
    bwd_ms = float(np.mean([b.elapsed_time(c) for _, b, c, _ in records]))
    pgd_ms = float(np.mean([c.elapsed_time(d) for _, _, c, d in records]))
    npx = B * S * S
    peak, peak_src = load_peaks()
    kern = {
        "edge_blend_fwd": {"ms": fwd_ms, "gbs": BYTES_FWD_PX * npx / fwd_ms / 1e6, "bytes_per_px": BYTES_FWD_PX},
        "edge_blend_bwd": {"ms": bwd_ms, "gbs": BYTES_BWD_PX * npx / bwd_ms / 1e6, "bytes_per_px": BYTES_BWD_PX},
        "pgd_linf_step": {"ms": pgd_ms, "gbs": BYTES_PGD_ELT * 3 * npx / pgd_ms / 1e6, "bytes_per_elt": BYTES_PGD_ELT},
    }
    for k in kern.values():
        k["frac_of_hbm_peak"] = k["gbs"] / peak
    share = {"edge_blend_fwd": 11 * fwd_ms, "edge_blend_bwd": 10 * bwd_ms, "pgd_linf_step": 10 * pgd_ms}
    dom = max(share, key=share.get)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": kern[dom]["gbs"] / peak, "traffic": ncu_traffic(dom, [B, 3, S, S], args.variant),
                "traffic_unit": "bytes per launch (dram read + write, ncu --set full capture in profiles/)",
                "algorithmic_bytes_per_launch": {"edge_blend_fwd": BYTES_FWD_PX, "edge_blend_bwd": BYTES_BWD_PX,
                                                 "pgd_linf_step": BYTES_PGD_ELT * 3}[dom] * npx,
                "peak_source": peak_src, "share_of_step": share[dom] / sum(share.values())}

    # ---- end to end through the public API with host buffers ---------------------------------
    e2e = e2e_api = None
    if not args.no_e2e:
        e2e = run_e2e(args, torch, dist, dev, world, rank, F_ee, p, (base, g_out), barrier)
        e2e_api = run_e2e_attack_api(args, torch, dist, dev, world, rank, core, attacks, canny, barrier)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, secs = cpu_hot_path(args.cpu_images, S, args.variant, repeats=1, min_seconds=args.cpu_seconds)
        n_cpu = cpu_hot_path.last_steps
        cpu = {"value": v, "unit": "images/s", "cores": cores, "kind": "port",
               "sample": "%d images of 3x%dx%d per step, %d full PGD-10 hot-path steps (%.1f s of CPU work, mean)"
                         % (args.cpu_images, S, S, n_cpu, secs * n_cpu)}
        live = cpu_live_reference(min(args.cpu_images, 256), S, args.variant)
        if live is not None:
            cpu["live_reference"] = live

    if args.sweep and rank == 0:
        run_sweep(torch, F_ee, canny, dev, peak, args.variant)

    named = extra = None
    if rank == 0 and world == 1:
        del x0, base, g_out, x_start, xa, xb, out, g_x, g_base
        torch.cuda.empty_cache()
        if not args.no_named_batch:
            named = run_named_batch(torch, F_ee, core, dev, peak)
        if not args.no_extra:
            extra = run_extra_configs(torch, core, attacks, dev)

    if rank == 0:
        line = {
            "metric": "edge-enhanced PGD-10 hot-path images/sec", "value": value, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args, B, world), host_cpus_bound_per_rank=numa_cpus),
            "clocks": clocks, "e2e": e2e, "e2e_attack_api": e2e_api, "gpu_launches": launches,
            "roofline": roofline, "kernels": kern, "cpu_baseline": cpu, "named_batch": named, "extra": extra,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _timed_pipeline(torch, dist, dev, world, barrier, step, steps, streams, main):
    """warm up 3 steps, then time `steps` steps that are issued on `streams` and only joined to `main` at the end
    (successive steps overlap like a prefetching input pipeline); max over ranks."""
    def join():
        for st in streams:
            ev = torch.cuda.Event()
            ev.record(st)
            main.wait_event(ev)

    for _ in range(3):
        step()
    join()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for st in streams:
        st.wait_event(e0)
    for _ in range(steps):
        step()
    join()
    e1.record(main)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def run_e2e(args, torch, dist, dev, world, rank, F_ee, p, resident, barrier):
    """The SAME step as `value` (10 x [fwd, bwd, PGD step] + final fwd, `base` and `g_out` resident like the CNN /
    FFT outputs they stand for) issued through the public functional API = the C-ABI entry points, but with HOST
    buffers: every step copies the clean batch from pinned host memory (H2D, 201 MB) and reads the step's result back.

    `e2e` (the contract's definition): the result read back is a METRIC -- the per-image mean of the blended adversarial
    image, B floats, computed by one torch reduction inside the timed region -- which is what a training step returns to the
    host (a loss); the adversarial batch itself stays on the device for the CNN.  `full_readback` (round 1's definition, kept
    for continuity): the whole adversarial batch goes back to the host too (another 201 MB per step, D2H).

    For each, two schedules are timed:
      pipeline : ONE whole-batch H2D on a dedicated copy-in stream, the 31 launches on a compute stream, the read-back on a
                 dedicated copy-out stream, triple buffered so that H2D(k+1), kernels(k) and D2H(k-1) overlap -- REPORTED;
      chunked  : the batch cut into --e2e-chunks chunks issued round-robin on --e2e-streams streams -- listed only (its
                 1024-image chunks are partly L2-resident between kernels, unlike the launches `value` times).
    Next to them the pure-CUDA copy floor of tools/copyfloor.cu for exactly that traffic (H2D only / both directions)."""
    B, S = args.batch, args.side
    shape = (B, 3, S, S)
    base, g_out = resident
    host_in = torch.rand(shape).pin_memory()
    host_out = torch.empty(shape).pin_memory()
    host_metric = torch.empty((B,)).pin_memory()
    main = torch.cuda.current_stream(dev)
    nbytes = B * 3 * S * S * 4
    steps = max(3, min(args.steps, 20))
    launches = [0]

    def hot_path(x0c, xa, xb, out, g_x, g_base, bs, go):
        cur, nxt = x0c, xa
        for _it in range(N_PGD):
            F_ee.edge_blend(cur, bs, p, W_BLEND, out=out)
            F_ee.edge_blend_backward(go, cur, bs, p, W_BLEND, g_x=g_x, g_base=g_base)
            F_ee.pgd_linf_step(cur, g_x, x0c, ALPHA, EPS, out=nxt)
            cur, nxt = nxt, (xb if nxt is xa else xa)
        F_ee.edge_blend(cur, bs, p, W_BLEND, out=out)
        launches[0] += 3 * N_PGD + 1
        return cur

    # ---- pipeline schedule
    s_up, s_run, s_dn = (torch.cuda.Stream(device=dev) for _ in range(3))
    NSLOT = 3                       # three stages (copy in, kernels, copy out) in flight need three buffer sets
    slots = [[torch.empty(shape, device=dev) for _ in range(3)] for _ in range(NSLOT)]  # per slot: x0, xa, xb
    metrics = [torch.empty((B,), device=dev) for _ in range(NSLOT)]
    scratch = [torch.empty(shape, device=dev) for _ in range(3)]                        # out, g_x, g_base (compute is serial)
    drained = [None] * NSLOT                                                            # event: the slot's D2H has finished
    counter = [0]

    def pipeline_step(with_kernels=True, full=False):
        k = counter[0]
        counter[0] += 1
        x0c, xa, xb = slots[k % NSLOT]
        if drained[k % NSLOT] is not None:
            s_up.wait_event(drained[k % NSLOT])
        with torch.cuda.stream(s_up):
            x0c.copy_(host_in, non_blocking=True)
            up = torch.cuda.Event()
            up.record(s_up)
        s_run.wait_event(up)
        with torch.cuda.stream(s_run):
            cur = hot_path(x0c, xa, xb, scratch[0], scratch[1], scratch[2], base, g_out) if with_kernels else x0c
            if not full:
                torch.mean((scratch[0] if with_kernels else x0c).view(B, -1), dim=1, out=metrics[k % NSLOT])
            ran = torch.cuda.Event()
            ran.record(s_run)
        s_dn.wait_event(ran)
        with torch.cuda.stream(s_dn):
            if full:
                host_out.copy_(cur, non_blocking=True)
            else:
                host_metric.copy_(metrics[k % NSLOT], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(s_dn)
        drained[k % NSLOT] = ev

    pstreams = [s_up, s_run, s_dn]
    T = lambda fn, streams: _timed_pipeline(torch, dist, dev, world, barrier, fn, steps, streams, main)
    ms_pipe = T(pipeline_step, pstreams)
    ms_pipe_copy = T(lambda: pipeline_step(False), pstreams)
    ms_pipe_full = T(lambda: pipeline_step(True, True), pstreams)
    ms_pipe_full_copy = T(lambda: pipeline_step(False, True), pstreams)
    del slots, scratch

    # ---- chunked schedule (round 1)
    n_chunks, N_STREAMS = max(1, args.e2e_chunks), max(1, args.e2e_streams)
    bounds = [(i * B // n_chunks, (i + 1) * B // n_chunks) for i in range(n_chunks)]
    bounds = [(lo, hi) for lo, hi in bounds if hi > lo]
    streams = [torch.cuda.Stream(device=dev) for _ in range(N_STREAMS)]
    bufs = [[torch.empty((hi - lo, 3, S, S), device=dev) for _ in range(6)] for lo, hi in bounds]   # x0, xa, xb, out, g_x, g_base
    dev_metric = torch.empty((B,), device=dev)

    def chunked_step(full=False):
        for i, (lo, hi) in enumerate(bounds):
            x0c, xa, xb, out, g_x, g_base = bufs[i]
            with torch.cuda.stream(streams[i % N_STREAMS]):
                x0c.copy_(host_in[lo:hi], non_blocking=True)
                cur = hot_path(x0c, xa, xb, out, g_x, g_base, base[lo:hi], g_out[lo:hi])
                if full:
                    host_out[lo:hi].copy_(cur, non_blocking=True)
                else:
                    torch.mean(out.view(hi - lo, -1), dim=1, out=dev_metric[lo:hi])
                    host_metric[lo:hi].copy_(dev_metric[lo:hi], non_blocking=True)

    ms_chunk = T(chunked_step, streams)
    ms_chunk_full = T(lambda: chunked_step(True), streams)
    del bufs
    torch.cuda.empty_cache()

    floor = copy_floor(torch, dist, dev, world, barrier, nbytes)
    floor_again = copy_floor(torch, dist, dev, world, barrier, nbytes)      # the box's contention is erratic for N > 2: show the spread
    label_chunk = "ms_per_step_chunked_%dx%d" % (len(bounds), N_STREAMS)

    def pack(ms_p, ms_c, ms_copy, d2h, floor_keys):
        # the reported number is the PIPELINE schedule: it launches the same 4096-image kernels as `value`.  The chunked
        # schedule works on 1024-image chunks (50 MB per tensor) that are partly L2-resident from one kernel to the next, which
        # is why it can be faster than the resident-input `value` itself; it is listed, not reported.
        ms = ms_p
        r = {"value": world * B * steps / (ms / 1e3), "unit": "images/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": d2h,
             "steps": steps, "ms_per_step": ms / steps, "schedule": "pipeline",
             label_chunk: ms_c / steps, "copies_only_ms_per_step": ms_copy / steps}
        try:
            fl = sorted(v["ms_per_step"] for f in (floor, floor_again) for k, v in f.items() if k.startswith(floor_keys))
            r["copy_floor_ms_per_step"] = fl[0]                    # the best the box did for this traffic
            r["copy_floor_ms_per_step_median"] = fl[len(fl) // 2]
            r["frac_of_copy_floor"] = fl[0] / (ms / steps)
            r["frac_of_copy_floor_median"] = fl[len(fl) // 2] / (ms / steps)
        except Exception:
            pass
        return r

    res = pack(ms_pipe, ms_chunk, ms_pipe_copy, B * 4, "h2d_only")
    res["result_read_back"] = "per-image mean of the blended adversarial image (B floats; a torch reduction inside the timed region)"
    res["full_readback"] = pack(ms_pipe_full, ms_chunk_full, ms_pipe_full_copy, nbytes, "both")
    res["full_readback"]["result_read_back"] = "the whole adversarial batch (round 1's definition of e2e)"
    res["copy_floor_pure_cuda"] = floor
    res["copy_floor_pure_cuda_second_pass"] = {k: v["ms_per_step"] for k, v in floor_again.items()} if "error" not in floor_again else floor_again
    res["api"] = ("functional.edge_blend / edge_blend_backward / pgd_linf_step (the C-ABI entry points), same 31-launch step as "
                  "`value`; clean batch H2D from pinned host memory every step; pipeline = one copy per direction per step on "
                  "dedicated copy streams, triple buffered")
    return res


def run_e2e_attack_api(args, torch, dist, dev, world, rank, core, attacks, canny, barrier):
    """attacks.PGD (the reference's call signature, utils/attacks.py:12) on a model whose front end is
    core.edge_enhance (base = x; the FFT low-pass is out of scope) and whose head is a per-channel mean -> 3-way
    cross-entropy (torch stand-in for the CNN: its kernels and the autograd bookkeeping are inside this number).
    Host buffers and chunking as in run_e2e."""
    n_chunks, N_STREAMS = max(1, args.e2e_chunks), max(1, args.e2e_streams)
    B, S = args.batch, args.side
    shape = (B, 3, S, S)
    host_in = torch.rand(shape).pin_memory()
    host_out = torch.empty(shape).pin_memory()
    targets = torch.randint(0, 3, (B,), device=dev)
    low = None if args.variant == "step125" else LOW
    bounds = [(i * B // n_chunks, (i + 1) * B // n_chunks) for i in range(n_chunks)]
    bounds = [(lo, hi) for lo, hi in bounds if hi > lo]
    streams = [torch.cuda.Stream(device=dev) for _ in range(N_STREAMS)]
    main = torch.cuda.current_stream(dev)

    def model(x):
        z = core.edge_enhance(x, x, canny, W_BLEND, low, HIGH, True)
        return z.flatten(2).mean(2) * 50.0

    class A:
        random = False
        epsilon = EPS

    def step():
        for i, (lo, hi) in enumerate(bounds):
            st = streams[i % N_STREAMS]
            with torch.cuda.stream(st):
                x = host_in[lo:hi].to(dev, non_blocking=True)
                x_adv = attacks.PGD(model, A, x, targets[lo:hi], N_PGD, ALPHA)
                with torch.no_grad():
                    model(x_adv)                               # the training forward on the adversarial batch
                host_out[lo:hi].copy_(x_adv, non_blocking=True)
                x.record_stream(st); x_adv.record_stream(st)

    steps = max(3, min(args.steps, 10))
    ms = _timed_pipeline(torch, dist, dev, world, barrier, step, steps, streams, main)
    nbytes = B * 3 * S * S * 4
    return {"value": world * B * steps / (ms / 1e3), "unit": "images/s",
            "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes, "steps": steps,
            "api": "attacks.PGD(model=edge_enhance front end + mean/CE head in torch, num_steps=10) + final forward; "
                   "%d chunks on %d streams, pinned host buffers" % (len(bounds), N_STREAMS)}


def run_sweep(torch, F_ee, canny, dev, peak, variant="step125"):
    """configs[4]: standalone kernel sweep (batch x side), printed on stderr as a table."""
    p = canny.params(None if variant == "step125" else LOW, HIGH, True)
    print("sweep (%s): B side | fwd GB/s (frac) | bwd GB/s (frac) | pgd GB/s (frac)" % variant, file=sys.stderr)
    for side in (32, 64, 224):
        for B in (64, 256, 1024, 4096):
            shape = (B, 3, side, side)
            x = torch.rand(shape, device=dev); base = torch.rand(shape, device=dev); g = torch.randn(shape, device=dev)
            o1, o2, o3 = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
            res = []
            for fn, nbytes in ((lambda: F_ee.edge_blend(x, base, p, 1.0, out=o1), 36.0),
                               (lambda: F_ee.edge_blend_backward(g, x, base, p, 1.0, g_x=o2, g_base=o3), 60.0),
                               (lambda: F_ee.pgd_linf_step(x, g, base, ALPHA, EPS, out=o1), 48.0)):
                for _ in range(3):
                    fn()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n = 20
                e0.record()
                for _ in range(n):
                    fn()
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / n
                gbs = nbytes * B * side * side / ms / 1e6
                res.append("%7.0f (%.2f) %6.1fus" % (gbs, gbs / peak, ms * 1e3))
            print("sweep: %5d %4d | %s | %s | %s" % (B, side, res[0], res[1], res[2]), file=sys.stderr)
            del x, base, g, o1, o2, o3


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
