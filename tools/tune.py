"""Sweep the row-strip height of the tiled edge kernels on the GPU and print algorithmic GB/s.
usage: python tools/tune.py [--shapes 4096x64,256x224,4096x32] [--variant step125]"""
import argparse
import contextlib
import io
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edge_enhancement_b200 import functional as F, _lib, core  # noqa: E402


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


def graph_timeit(fn, reps=20):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    return timeit(g.replay, 10) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="4096x64,512x224,8192x32,16384x28x1")
    ap.add_argument("--variant", default="step125")
    ap.add_argument("--ths", default="0,4,8,16,32,64,112,224")
    ap.add_argument("--channels-last", action="store_true")
    ap.add_argument("--staging", type=int, default=0, help="ee_set_tuning staging knob (8 / 9 = always / never row-streaming)")
    ap.add_argument("--graph", action="store_true", help="time CUDA-graph replays of 20 launches (small batches: no host overhead)")
    args = ap.parse_args()
    L = _lib.load()
    with contextlib.redirect_stdout(io.StringIO()):
        f = {"step125": core.CannyFilter_step125_1, "canny": core.CannyFilter, "bpda": core.CannyFilter_BPDA}[args.variant]()
    p = f.params(None if args.variant == "step125" else 38 / 255, 76 / 255, True)
    for spec in args.shapes.split(","):
        parts = [int(v) for v in spec.split("x")]
        B, S = parts[0], parts[1]
        C = parts[2] if len(parts) > 2 else 3
        x = torch.rand(B, C, S, S, device="cuda"); base = torch.rand_like(x) * 1.1 - 0.1; g = torch.randn_like(x)
        if args.channels_last:
            x, base, g = (t.contiguous(memory_format=torch.channels_last) for t in (x, base, g))
        o1, o2, o3 = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
        npx = B * S * S
        for th in [int(t) for t in args.ths.split(",")]:
            if th > S:
                continue
            L.ee_set_tuning(th, th, args.staging)
            tm = graph_timeit if args.graph else timeit
            tf = tm(lambda: F.edge_blend(x, base, p, 1.0, out=o1))
            tb = tm(lambda: F.edge_blend_backward(g, x, base, p, 1.0, g_x=o2, g_base=o3))
            print("%-8s B=%5d C=%d side=%3d TH=%3d | fwd %7.1f us %6.0f GB/s | bwd %7.1f us %6.0f GB/s"
                  % (args.variant, B, C, S, th, tf * 1e3, 12.0 * C * npx / tf / 1e6, tb * 1e3, 20.0 * C * npx / tb / 1e6), flush=True)
        L.ee_set_tuning(0, 0, 0)


if __name__ == "__main__":
    main()
