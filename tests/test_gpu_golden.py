"""-m gpu: the CUDA path against the fixtures produced by the UNMODIFIED reference
(tests/golden/*.npz, oracle/make_golden.py), through the drop-in Python surface.

Tolerances (BASELINE.json north_star): threshold masks exact; blended images 1e-5 relative; gradients
1e-5 of max|g| where the reference gradient is finite; attack steps bit-exact; full PGD-10 runs may
differ from the reference only on elements whose gradient sign is ambiguous."""
import contextlib
import glob
import io
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from edge_enhancement_b200 import attacks, core, functional as F_ee   # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EDGE_FILES = sorted(glob.glob(os.path.join(GOLD, "edge_*.npz")))
DEV = "cuda:0"
CLS = {"step125": core.CannyFilter_step125_1, "canny": core.CannyFilter, "bpda": core.CannyFilter_BPDA}


def _opt(s):
    return None if s == "None" else float(s)


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def cu(a, grad=False):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return t.requires_grad_() if grad else t


@pytest.mark.parametrize("fused", [False, True], ids=["module+torch_blend", "fused_edge_enhance"])
@pytest.mark.parametrize("path", EDGE_FILES, ids=lambda p: os.path.basename(p)[5:-4])
def test_dropin_modules_match_reference_fixture(path, fused):
    z = np.load(path)
    variant, alpha, sigma, low, high, hyst, w = [str(v) for v in z["meta"]]
    alpha, sigma, low, high, hyst, w = float(alpha), float(sigma), _opt(low), _opt(high), bool(int(hyst)), float(w)
    f = quiet(CLS[variant], sigma=sigma, use_cuda=False, alpha=alpha)
    x, base = cu(z["x"], True), cu(z["base"], True)
    if fused:
        out = core.edge_enhance(x, base, f, w, low, high, hyst)
        edge = f(x.detach(), low_threshold=low, high_threshold=high, hysteresis=hyst)
    else:
        edge = f(x, low_threshold=low, high_threshold=high, hysteresis=hyst)        # reference call signature
        out = torch.clamp(base + w * edge, 0.0, 1.0)                                # resnet_EE.py:189-191
    out.backward(cu(z["g_out"]))
    binary = not (low is None or (variant == "bpda" and high is None))
    e = edge.detach().cpu().numpy()
    if binary:
        assert np.array_equal(e, z["edge"]), "%d mask pixels differ" % (e != z["edge"]).sum()
    else:
        np.testing.assert_allclose(e, z["edge"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(out.detach().cpu().numpy(), z["out"], rtol=1e-5, atol=1e-7)
    assert np.array_equal(base.grad.cpu().numpy(), z["g_base"])
    ref = z["g_x"]
    fin = np.isfinite(ref)
    g = x.grad.cpu().numpy()
    assert np.isfinite(g).all()
    assert np.abs(g - ref)[fin].max() <= 1e-5 * np.abs(ref[fin]).max()


def test_attack_step_fixtures():
    z = np.load(os.path.join(GOLD, "attack_steps.npz"))
    eps, a = 16 / 255, 2 / 255
    x, g, x0 = cu(z["x"]), cu(z["g"]), cu(z["x0"])
    eq = lambda t, n: np.array_equal(t.cpu().numpy(), z[n], equal_nan=True)
    assert eq(F_ee.pgd_linf_step(x, g, x0, a, eps), "pgd")
    assert eq(F_ee.pgd_linf_step(x, g, x0, -a, eps), "tpgd")
    assert eq(F_ee.fgsm_step(x, g, 0.007), "fgsm")
    d = (x - x0).contiguous()
    adv = F_ee.free_at_step_(d, g, x0, 4 / 255, 4 / 255)
    assert eq(d, "free_delta") and eq(adv, "free_adv")
    assert eq(F_ee.cw_linf_step(x, g, x0, cu(z["cw_min"]), cu(z["cw_max"]), 0.00392, 0.02), "cw")
    np.testing.assert_allclose(F_ee.pgd_l2_step(x, cu(z["g2"]), x0, 0.5, 0.02).cpu().numpy(), z["l2"], rtol=1e-5, atol=1e-6)


class TinyEENet(torch.nn.Module):
    """Same model as oracle/make_golden.py::TinyEENet with the drop-in filter as front end."""

    def __init__(self, canny, low, high, w, C, H, W, n_class, seed):
        super().__init__()
        self.canny, self.low, self.high, self.w = canny, low, high, w
        r = np.random.default_rng(seed)
        self.weight = torch.from_numpy(r.standard_normal((n_class, C * H * W)).astype(np.float32) * 0.05).to(DEV)
        self.grads = []

    def forward(self, x):
        if x.requires_grad:
            x.register_hook(lambda g: self.grads.append(g.detach().clone()))
        e = self.canny(x, low_threshold=self.low, high_threshold=self.high, hysteresis=True)
        z = torch.clamp(x + self.w * e, 0.0, 1.0)
        return z.reshape(z.shape[0], -1) @ self.weight.t()


@pytest.mark.parametrize("variant", ["step125", "canny"])
def test_full_pgd10_against_reference_run(variant):
    """utils.attacks.PGD (reference, CPU, fixture) vs attacks.PGD (drop-in, GPU) on the same model."""
    z = np.load(os.path.join(GOLD, "pgd10_%s.npz" % variant))
    x, y = cu(z["x"]), torch.from_numpy(z["y"]).to(DEV)
    B, C, H, W = x.shape
    torch.backends.cuda.matmul.allow_tf32 = False
    model = TinyEENet(quiet(CLS[variant], use_cuda=False, alpha=0.0), 38 / 255, 76 / 255, 1.0, C, H, W,
                      int(z["n_class"]), int(z["head_seed"]))

    class Args:
        random = False
        epsilon = 16 / 255
    x_adv = attacks.PGD(model, Args, x, y, 10, 2 / 255)
    got, want = x_adv.cpu().numpy(), z["x_adv"]
    assert got.min() >= 0 and got.max() <= 1 and np.abs(got - z["x"]).max() <= 16 / 255 + 1e-6
    bad = np.abs(got - want) > 1e-6
    # every differing element must have had an ambiguous gradient sign at some step of our own run
    gmax = max(float(g.abs().max()) for g in model.grads)
    ambiguous = np.zeros_like(bad)
    for g in model.grads:
        ambiguous |= (g.abs().cpu().numpy() < 1e-5 * gmax)
    assert bad.mean() < 0.01, "%.3f %% of the adversarial example differs" % (100 * bad.mean())
    unexplained = bad & ~ambiguous
    assert unexplained.sum() <= 0.002 * bad.size, "%d differing elements are not sign-ambiguous" % unexplained.sum()


def test_attack_extras_fixture():
    """random start and AVmixup vertex / mix: CUDA kernels vs the reference's torch expressions (bit-exact)"""
    z = np.load(os.path.join(GOLD, "attack_extras.npz"))
    start = F_ee.add_clamp(cu(z["x0"]), cu(z["noise"]))
    assert np.array_equal(start.cpu().numpy(), z["start"])
    w = torch.from_numpy(z["weight"]).to(DEV)
    mixed = F_ee.avmixup_mix(cu(z["x"]), cu(z["x0"]), w, float(z["gamma"]))
    assert np.array_equal(mixed.cpu().numpy(), z["mixed"])
    # the same numbers through the drop-in attack helper: seeded torch generator -> same noise as the reference draw
    torch.manual_seed(601)
    x0 = cu(z["x0"])
    ref_noise = torch.zeros_like(x0).uniform_(-float(z["eps"]), float(z["eps"]))
    torch.manual_seed(601)
    got = attacks._random_start(x0, float(z["eps"]))
    assert torch.equal(got, torch.clamp(x0 + ref_noise, 0, 1))
