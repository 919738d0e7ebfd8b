"""Module-level (zero-diff drop-in) edge kernels: edge = CannyFilter*(img) without the fused blend.
fwd: read x (4C) + write edge (4) B/px; bwd: read g_edge (4) + x (4C) + write g_x (4C) B/px.
usage: python tools/tune_module_level.py"""
import contextlib, io, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edge_enhancement_b200 import functional as F, core  # noqa: E402
from tools.tune import timeit  # noqa: E402

for variant in ("step125", "canny", "bpda"):
    with contextlib.redirect_stdout(io.StringIO()):
        f = {"step125": core.CannyFilter_step125_1, "canny": core.CannyFilter, "bpda": core.CannyFilter_BPDA}[variant]()
    p = f.params(None if variant == "step125" else 38 / 255, 76 / 255, True)
    for B, S, C in ((4096, 64, 3), (512, 224, 3), (16384, 28, 1)):
        x = torch.rand(B, C, S, S, device="cuda"); ge = torch.randn(B, 1, S, S, device="cuda")
        npx = B * S * S
        tf = timeit(lambda: F.edge_map(x, p))
        tb = timeit(lambda: F.edge_map_backward(ge, x, p))
        print("%-8s module-level B=%5d C=%d side=%3d | fwd %7.1f us %6.0f GB/s | bwd %7.1f us %6.0f GB/s"
              % (variant, B, C, S, tf * 1e3, (4 * C + 4) * npx / tf / 1e6, tb * 1e3, (8 * C + 4) * npx / tb / 1e6), flush=True)
