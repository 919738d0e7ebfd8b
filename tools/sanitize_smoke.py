"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py
Shapes are tiny (the tools slow kernels down 10-100x) but cover: whole-image (HT) kernels 64/32/28 px, strips, chunk-aligned
tiles with and without TMA staging, the cluster kernel, Canny / BPDA forward + backward, every elementwise kernel."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edge_enhancement_b200 import functional as F, _lib  # noqa: E402
from oracle import oracle as O  # noqa: E402

L = _lib.load()
dev = "cuda:0"
g3 = O.gaussian3()
torch.manual_seed(0)
n = 0
for variant, low in (("step125", None), ("canny", 38 / 255), ("bpda", 38 / 255)):
    p = F.make_params(variant, g3, 0.0, low, 76 / 255, True)
    for shape, stagings in (((2, 3, 64, 64), (0, 6, 1)), ((2, 3, 32, 32), (0,)), ((3, 1, 28, 28), (0,)), ((1, 3, 224, 224), (0, 3, 5, 7)),
                            ((1, 3, 40, 300), (0, 3)), ((1, 3, 17, 23), (0,))):
        x = torch.rand(shape, device=dev); base = torch.rand(shape, device=dev) * 1.1 - 0.1; g = torch.randn(shape, device=dev)
        for st in stagings:
            L.ee_set_tuning(0, 0, st)
            out = F.edge_blend(x, base, p, 1.0)
            gx, gb = F.edge_blend_backward(g, x, base, p, 1.0)
            e = F.edge_map(x, p)
            ge = F.edge_map_backward(g[:, :1].contiguous(), x, p)
            n += 4
L.ee_set_tuning(0, 0, 0)
x = torch.rand(4, 3, 16, 16, device=dev); g = torch.randn_like(x); x0 = torch.rand_like(x)
F.pgd_linf_step(x, g, x0, 2 / 255, 16 / 255); F.fgsm_step(x, g, 0.01); F.cw_linf_step(x, g, x0, x0 - 0.03, x0 + 0.03, 0.004, 0.02)
F.pgd_l2_step(x, g, x0, 0.5, 0.02); F.add_clamp(x, g * 0.01); F.free_at_step_(torch.zeros_like(x), g, x0, 4 / 255, 4 / 255)
F.avmixup_mix(x, x0, torch.rand(4, dtype=torch.float64, device=dev), 2.0)
stripe = torch.sign(torch.rand(4, 3, 16, device=dev) - 0.5); table = torch.tensor([[3., 5., .1, -.1, .1]], device=dev)
F.add_square(x, stripe, table, 0.05); F.add_square_backward(g, x, stripe, table, 0.05)
torch.cuda.synchronize()
print("sanitize_smoke: %d edge calls + elementwise kernels done" % n)
