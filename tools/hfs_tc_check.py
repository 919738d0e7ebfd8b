"""A/B of the two 64 px HighFreqSuppress kernels (ee_hfs_f32: FFMA, ee_hfs_tc_f32: tcgen05): error against float64 and time."""
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
    return e0.elapsed_time(e1) / n * 1e3


def main():
    L = _lib.load()
    torch.manual_seed(0)
    HF = core.HighFreqSuppress(64, 64, 8, impl='torch_fft')
    for planes in (3, 1000, 12288):
        x = torch.rand(planes, 64, 64, device="cuda")
        add = torch.randn(planes, 64, 64, device="cuda")
        for use_add in (False, True):
            ref = F.hfs(x, 8, add=add if use_add else None)
            got = F.hfs(x, 8, add=add if use_add else None, impl='tcgen05')
            torch.cuda.synchronize()
            exact = HF._fft_forward(x.double()) + (add.double() if use_add else 0.0)
            err = (got - ref).abs().max().item()
            e_tc = (got.double() - exact).abs().max().item()
            e_ff = (ref.double() - exact).abs().max().item()
            print("planes %6d add %d: max|tc - ffma| = %.3e  |tc - f64| = %.3e  |ffma - f64| = %.3e  (max|y| %.3f)  nan %d"
                  % (planes, use_add, err, e_tc, e_ff, ref.abs().max().item(), int(torch.isnan(got).sum())), flush=True)
    for planes in (768, 12288):
        x = torch.rand(planes, 64, 64, device="cuda")
        y = torch.empty_like(x)
        for impl in ('native', 'tcgen05', 'native', 'tcgen05'):
            print("planes %5d %-8s: %.1f us" % (planes, impl, timeit(lambda: F.hfs(x, 8, out=y, impl=impl))), flush=True)


if __name__ == "__main__":
    main()
