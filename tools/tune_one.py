"""Time fwd / bwd of one shape under a given ee_set_tuning(th_fwd, th_bwd, staging).
usage: python tools/tune_one.py 512x224 [variant] [th] [staging]"""
import contextlib, io, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edge_enhancement_b200 import functional as F, _lib, core  # noqa: E402
from tools.tune import timeit  # noqa: E402

spec = sys.argv[1]
variant = sys.argv[2] if len(sys.argv) > 2 else "step125"
th = int(sys.argv[3]) if len(sys.argv) > 3 else 0
staging = int(sys.argv[4]) if len(sys.argv) > 4 else 0
parts = [int(v) for v in spec.split("x")]
B, S = parts[0], parts[1]
C = parts[2] if len(parts) > 2 else 3
L = _lib.load()
with contextlib.redirect_stdout(io.StringIO()):
    f = {"step125": core.CannyFilter_step125_1, "canny": core.CannyFilter, "bpda": core.CannyFilter_BPDA}[variant]()
p = f.params(None if variant == "step125" else 38 / 255, 76 / 255, True)
x = torch.rand(B, C, S, S, device="cuda"); base = torch.rand_like(x) * 1.1 - 0.1; g = torch.randn_like(x)
o1, o2, o3 = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
L.ee_set_tuning(th, th, staging)
npx = B * S * S
tf = timeit(lambda: F.edge_blend(x, base, p, 1.0, out=o1))
tb = timeit(lambda: F.edge_blend_backward(g, x, base, p, 1.0, g_x=o2, g_base=o3))
print("%-8s B=%5d C=%d side=%3d TH=%3d staging=%d | fwd %7.1f us %6.0f GB/s | bwd %7.1f us %6.0f GB/s"
      % (variant, B, C, S, th, staging, tf * 1e3, 12.0 * C * npx / tf / 1e6, tb * 1e3, 20.0 * C * npx / tb / 1e6), flush=True)
