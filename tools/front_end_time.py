"""Forward + backward time of the WHOLE edge-enhancement front end of the *_EE models (low-pass, edge filter, blend,
clamp; resnet_EE.py:176-191) at a given batch: this package's fused node (core.EdgeEnhance: 3 kernels per direction)
vs the reference-style eager composition (tools/train_throughput.py:EagerFront, stock torch ops + torch.fft).
usage: python tools/front_end_time.py [B] [side]"""
import contextlib, io, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edge_enhancement_b200 import core  # noqa: E402
from tools.tune import timeit  # noqa: E402
from tools.train_throughput import EagerFront  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
S = int(sys.argv[2]) if len(sys.argv) > 2 else 64
r = {28: 4, 64: 8, 224: 16}[S]
C = 1 if S == 28 else 3
dev = "cuda:0"
x = torch.rand(B, C, S, S, device=dev)
g = torch.randn_like(x)
with contextlib.redirect_stdout(io.StringIO()):
    ours = core.EdgeEnhance(cize=S, r=r, w=1.0, low=38.0, high=76.0, alpha=0.0, sigma=1, type_canny='CannyFilter_step125_1')
fronts = {"ours (fused node)": ours, "eager_clean (no dead host allocations)": EagerFront(S, r, 1.0, 76 / 255, dev, faithful=False)}
if S == 64:       # the same node on the tensor-core low-pass (ee_hfs_tc_f32)
    with contextlib.redirect_stdout(io.StringIO()):
        fronts["ours (fused node, hfs_impl='tcgen05')"] = core.EdgeEnhance(cize=S, r=r, w=1.0, low=38.0, high=76.0, alpha=0.0, sigma=1,
                                                                            type_canny='CannyFilter_step125_1', hfs_impl='tcgen05')
if B * C * S * S <= 64 * 3 * 224 * 224 * 8:
    fronts["eager (reference-style, incl. per-call host scratch + H2D)"] = EagerFront(S, r, 1.0, 76 / 255, dev, faithful=True)
res = {}
for name, f in fronts.items():
    xr = x.clone().requires_grad_()
    def fb():
        y = f(xr); y.backward(g); xr.grad = None
    with torch.no_grad():
        tf = timeit(lambda: f(x), n=10)
    tfb = timeit(fb, n=10)
    res[name] = {"fwd_us": tf * 1e3, "fwd_bwd_us": tfb * 1e3}
print(json.dumps({"shape": [B, C, S, S], "front_ends": res,
                  "speedup_fwd_bwd_vs_eager_clean": res["eager_clean (no dead host allocations)"]["fwd_bwd_us"] / res["ours (fused node)"]["fwd_bwd_us"]}))
