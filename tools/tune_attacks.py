"""Time the elementwise / per-sample attack kernels and the per-call host overhead at the reference's batch sizes.
usage: python tools/tune_attacks.py [--skip-small]"""
import argparse
import contextlib
import io
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edge_enhancement_b200 import functional as F, _lib, core  # noqa: E402

PEAK = 6453.1


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def graph_time(fn, reps=20, replays=10):
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
    return timeit(g.replay, replays) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-small", action="store_true")
    args = ap.parse_args()
    L = _lib.load()
    print("== PGD-L2 step (16 B/elt): one-pass cluster kernel vs three-pass kernel")
    for B, C, S in ((4096, 3, 64), (512, 3, 224), (16384, 1, 28), (8192, 3, 32), (1024, 3, 128), (256, 3, 64), (32, 3, 224)):
        x = torch.rand(B, C, S, S, device="cuda"); g = torch.randn_like(x); x0 = torch.rand_like(x)
        n = x.numel()
        res = []
        for staging in (0, 1):
            L.ee_set_tuning(0, 0, staging)
            t = timeit(lambda: F.pgd_l2_step(x, g, x0, 0.003, 0.047))
            res.append("%8.1f us %6.0f GB/s (%.2f)" % (t * 1e3, 16.0 * n / t / 1e6, 16.0 * n / t / 1e6 / PEAK))
        L.ee_set_tuning(0, 0, 0)
        tl = timeit(lambda: F.pgd_linf_step(x, g, x0, 2 / 255, 16 / 255))
        print("l2 B=%5d C=%d side=%3d | cluster %s | three-pass %s | linf step %7.1f us %6.0f GB/s" %
              (B, C, S, res[0], res[1], tl * 1e3, 16.0 * n / tl / 1e6), flush=True)
        del x, g, x0
    print("== with_gf blend (fwd 28 B/px at C=3: edge + base + out; bwd 56 B/px)")
    for B, C, S in ((4096, 3, 64), (512, 3, 224)):
        edge = (torch.rand(B, 1, S, S, device="cuda") > 0.8).float(); base = torch.rand(B, C, S, S, device="cuda"); g = torch.randn_like(base)
        gs = core.get_gaussian_kernel(3, 0., 1.)
        tf = timeit(lambda: F.gf_blend(edge, base, gs, 1.0))
        tb = timeit(lambda: F.gf_blend_backward(g, edge, base, gs, 1.0))
        npx = B * S * S
        print("gf B=%5d side=%3d | fwd %7.1f us %6.0f GB/s | bwd %7.1f us %6.0f GB/s" %
              (B, S, tf * 1e3, (8 * C + 4) * npx / tf / 1e6, tb * 1e3, (12 * C + 8) * npx / tb / 1e6), flush=True)
    if args.skip_small:
        return
    print("== per-call cost at the reference's batch sizes: eager API call vs CUDA-graph replay (us per launch)")
    with contextlib.redirect_stdout(io.StringIO()):
        filt = {"step125": core.CannyFilter_step125_1(), "canny": core.CannyFilter(alpha=0.3)}
    for name, B, C, S, variant, low, high in (("T 256x3x64x64", 256, 3, 64, "step125", None, 76 / 255), ("M 128x1x28x28", 128, 1, 28, "canny", 25 / 255, 51 / 255),
                                              ("I 32x3x224x224", 32, 3, 224, "step125", None, 76 / 255), ("B=64 3x64x64", 64, 3, 64, "step125", None, 76 / 255)):
        p = filt[variant].params(low, high, True)
        x = torch.rand(B, C, S, S, device="cuda"); base = torch.rand_like(x) * 1.1 - 0.1; g = torch.randn_like(x); x0 = torch.rand_like(x)
        o1, o2, o3, o4 = (torch.empty_like(x) for _ in range(4))
        npx = B * S * S
        ops = (("fwd", lambda: F.edge_blend(x, base, p, 1.0, out=o1), 12.0 * C * npx),
               ("bwd", lambda: F.edge_blend_backward(g, x, base, p, 1.0, g_x=o2, g_base=o3), 20.0 * C * npx),
               ("step", lambda: F.pgd_linf_step(x, g, x0, 2 / 255, 16 / 255, out=o4), 16.0 * C * npx),
               ("iter(1 call)", lambda: F.pgd_iteration(x, base, g, x0, p, 1.0, 2 / 255, 16 / 255, out=o1, g_x=o2, g_base=o3, x_next=o4), 48.0 * C * npx))
        for op, fn, nbytes in ops:
            te = timeit(fn, 200)
            # host-only cost of the call: launches issued without waiting for the GPU
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(200):
                fn()
            host = (time.perf_counter() - t0) / 200
            torch.cuda.synchronize()
            tg = graph_time(fn)
            print("%-15s %-12s | eager %6.2f us (host issue %5.2f us) | graph %6.2f us = %6.0f GB/s (%.2f of HBM peak; L2-resident)" %
                  (name, op, te * 1e3, host * 1e6, tg * 1e3, nbytes / tg / 1e6, nbytes / tg / 1e6 / PEAK), flush=True)


if __name__ == "__main__":
    main()
