"""Launch the fused fwd / bwd / pgd kernels a few times on one shape (for ncu captures).
usage: python tools/prof_one.py --shape 4096x64 [--th-fwd N --th-bwd N --staging S --variant step125 --iters 3]"""
import argparse
import contextlib
import io
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edge_enhancement_b200 import functional as F, _lib, core  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="4096x64")
ap.add_argument("--variant", default="step125")
ap.add_argument("--th-fwd", type=int, default=0)
ap.add_argument("--th-bwd", type=int, default=0)
ap.add_argument("--staging", type=int, default=0)
ap.add_argument("--iters", type=int, default=3)
args = ap.parse_args()
parts = [int(v) for v in args.shape.split("x")]
B, S = parts[0], parts[1]
C = parts[2] if len(parts) > 2 else 3
_lib.load().ee_set_tuning(args.th_fwd, args.th_bwd, args.staging)
with contextlib.redirect_stdout(io.StringIO()):
    f = {"step125": core.CannyFilter_step125_1, "canny": core.CannyFilter, "bpda": core.CannyFilter_BPDA}[args.variant]()
p = f.params(None if args.variant == "step125" else 38 / 255, 76 / 255, True)
x = torch.rand(B, C, S, S, device="cuda"); base = torch.rand_like(x) * 1.1 - 0.1; g = torch.randn_like(x)
o1, o2, o3 = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
for _ in range(args.iters):
    F.edge_blend(x, base, p, 1.0, out=o1)
    F.edge_blend_backward(g, x, base, p, 1.0, g_x=o2, g_base=o3)
    F.pgd_linf_step(x, o2, base, 2 / 255, 16 / 255, out=o1)
torch.cuda.synchronize()
print("ok", float(o1.sum()), float(o2.abs().sum()))
