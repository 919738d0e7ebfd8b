"""-m gpu: round-2 additions, through the C ABI.

* reference-NaN compatible backward (EE_FLAG_NAN_COMPAT) vs the reference fixtures and the oracle;
* stand-alone STE Functions vs the reference fixture;
* with_gf=True blend (ee_gf_blend_{fwd,bwd}_f32) vs the reference fixture and the oracle;
* teacher-forced replay of the reference's own PGD-10 runs (mask exact, gradient 1e-5, step bit-exact);
* one-pass cluster PGD-L2 step and the single-call PGD iteration vs the oracle (bit-exact).
"""
import contextlib
import glob
import io
import os

import numpy as np
import pytest
import torch

from tests import common as T

pytestmark = pytest.mark.gpu

from edge_enhancement_b200 import _lib, attacks, core, functional as F_ee   # noqa: E402
from oracle import oracle as O                                              # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEV = "cuda:0"
GAUSS = O.gaussian3(0.0, 1.0)
CLS = {"step125": core.CannyFilter_step125_1, "canny": core.CannyFilter, "bpda": core.CannyFilter_BPDA}


def cu(a, grad=False):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return t.requires_grad_() if grad else t


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def _opt(s):
    return None if s == "None" else float(s)


def same(got, want):
    got = got.detach().cpu().numpy() if isinstance(got, torch.Tensor) else got
    return np.array_equal(got, want, equal_nan=True)


@pytest.fixture(autouse=True)
def _reset_tuning():
    _lib.load().ee_set_tuning(0, 0, 0)
    yield
    _lib.load().ee_set_tuning(0, 0, 0)


# ---------------------------------------------------------------------------------------------
# NaN-compatible backward
# ---------------------------------------------------------------------------------------------
NAN_FILES = sorted(glob.glob(os.path.join(GOLD, "nan_*.npz")))


@pytest.mark.parametrize("fused", [False, True], ids=["module", "fused"])
@pytest.mark.parametrize("path", NAN_FILES, ids=lambda p: os.path.basename(p)[4:-4])
def test_nan_compat_matches_reference_fixture(path, fused):
    """module.nan_compat = True: the reference's NaN set exactly (flat regions), finite entries within 1e-5."""
    assert len(NAN_FILES) >= 4
    z = np.load(path)
    variant, alpha, sigma, low, high, hyst, w = [str(v) for v in z["meta"]]
    alpha, low, high, w = float(alpha), _opt(low), _opt(high), float(w)
    f = quiet(CLS[variant], use_cuda=False, alpha=alpha)
    f.nan_compat = True
    x, base = cu(z["x"], True), cu(z["base"], True)
    if fused:
        out = core.edge_enhance(x, base, f, w, low, high, True)
    else:
        out = torch.clamp(base + w * f(x, low_threshold=low, high_threshold=high, hysteresis=True), 0.0, 1.0)
    out.backward(cu(z["g_out"]))
    g, ref = x.grad.cpu().numpy(), z["g_x"]
    assert np.array_equal(np.isnan(g), np.isnan(ref)), "%d entries differ in NaN-ness" % (np.isnan(g) != np.isnan(ref)).sum()
    fin = np.isfinite(ref)
    assert np.abs(g - ref)[fin].max() <= 1e-5 * np.abs(ref[fin]).max()
    assert np.array_equal(base.grad.cpu().numpy(), z["g_base"])
    # and bit for bit what the oracle says
    po = O.make_params(variant, alpha=alpha, low=low, high=high, hysteresis=True, nan_compat=True)
    o_gx, _ = O.edge_blend_bwd(z["g_out"], z["x"], z["base"], po, w)
    if fused:                                   # (the unfused blend sums the channels in torch's order)
        assert same(g, o_gx)
    # the default stays finite
    f.nan_compat = False
    x2 = cu(z["x"], True)
    core.edge_enhance(x2, cu(z["base"]), f, w, low, high, True).backward(cu(z["g_out"]))
    assert torch.isfinite(x2.grad).all()


@pytest.mark.parametrize("shape", [(2, 3, 64, 64), (3, 1, 28, 28), (1, 3, 224, 224), (2, 3, 17, 23), (2, 3, 40, 44)], ids=str)
@pytest.mark.parametrize("variant", ["step125", "canny", "bpda"])
def test_nan_compat_vs_oracle(variant, shape):
    x, base, g_out, g_edge = T.make_inputs(31, *shape, kind="sparse")
    x[:, :, : shape[2] // 3, :] = 0.25              # a flat band touching three image borders
    low = None if variant == "step125" else T.LOW
    pc = F_ee.make_params(variant, GAUSS, 0.0, low, T.HIGH, True, nan_compat=True)
    po = O.make_params(variant, alpha=0.0, low=low, high=T.HIGH, hysteresis=True, nan_compat=True)
    g_x, g_base = F_ee.edge_blend_backward(cu(g_out), cu(x), cu(base), pc, 1.0)
    o_gx, o_gb = O.edge_blend_bwd(g_out, x, base, po, 1.0)
    assert np.isnan(o_gx).any()
    assert same(g_x, o_gx) and same(g_base, o_gb)
    assert same(F_ee.edge_map_backward(cu(g_edge), cu(x), pc), O.edge_bwd(g_edge, x, po))
    # channels_last input: converted, same numbers
    xc = cu(x).contiguous(memory_format=torch.channels_last)
    g2, _ = F_ee.edge_blend_backward(cu(g_out), xc, cu(base), pc, 1.0)
    assert same(g2.contiguous(), o_gx)
    # forward is unaffected by the flag
    assert same(F_ee.edge_blend(cu(x), cu(base), pc, 1.0), O.edge_blend_fwd(x, base, po, 1.0))


def test_pgd_freezes_nan_pixels_like_the_reference():
    """torch.sign(NaN) = 0: with nan_compat the fused PGD step leaves the flat-region pixels where they are."""
    z = np.load(NAN_FILES[0])
    g = cu(z["g_x"])
    x = cu(z["x"])
    out = F_ee.pgd_linf_step(x, g, x, 2 / 255, 16 / 255)
    nan = torch.isnan(g)
    assert nan.any() and torch.equal(out[nan], x[nan]) and not torch.isnan(out).any()


# ---------------------------------------------------------------------------------------------
# stand-alone STE Functions vs the reference fixture
# ---------------------------------------------------------------------------------------------
def test_ste_functions_match_reference_fixture():
    z = np.load(os.path.join(GOLD, "ste_functions.npz"))
    v, g = z["v"], z["g"]
    for tag in ("pos", "neg", "zero"):
        t = cu(v, True)
        y = core.To_compare.apply(t, torch.tensor(float(z["cmp_%s_thr" % tag])))
        y.backward(cu(g))
        assert same(y, z["cmp_%s_fwd" % tag]) and same(t.grad, z["cmp_%s_bwd" % tag]), tag
    t = cu(v, True); y = core.To_eq.apply(t); y.backward(cu(g))
    assert same(y, z["eq_fwd"]) and same(t.grad, z["eq_bwd"])
    t = cu(v, True); y = core.BinaryConnectDeterministic.apply(t); y.backward(cu(g))
    assert same(y, z["bcd_fwd"]) and same(t.grad, z["bcd_bwd"])
    assert same(core.safeSign(cu(v)), z["safe_sign"])


# ---------------------------------------------------------------------------------------------
# with_gf=True
# ---------------------------------------------------------------------------------------------
def test_gf_blend_matches_reference_fixture():
    z = np.load(os.path.join(GOLD, "gf_blend.npz"))
    w, high = float(z["w"]), float(z["high"])
    f = quiet(core.CannyFilter_step125_1, use_cuda=False, alpha=0.0)
    x, base = cu(z["x"], True), cu(z["base"], True)
    out = core.edge_enhance(x, base, f, w, 38 / 255, high, True, with_gf=True)
    out.backward(cu(z["g_out"]))
    np.testing.assert_allclose(out.detach().cpu().numpy(), z["out"], rtol=1e-5, atol=1e-7)
    assert np.array_equal(base.grad.cpu().numpy(), z["g_base"])
    fin = np.isfinite(z["g_x"])
    assert np.abs(x.grad.cpu().numpy() - z["g_x"])[fin].max() <= 1e-5 * np.abs(z["g_x"][fin]).max()
    # module form
    m = quiet(core.EdgeEnhance, cize=24, r=4, w=w, low=38.0, high=76.0, type_canny='CannyFilter_step125_1', hfs=False, with_gf=True)
    assert torch.equal(m(x.detach()), core.edge_enhance(x.detach(), x.detach(), f, w, 38 / 255, high, True, with_gf=True))


@pytest.mark.parametrize("shape", [(2, 3, 24, 40), (3, 1, 28, 28), (1, 3, 17, 23), (2, 3, 64, 64), (1, 2, 50, 70)], ids=str)
def test_gf_kernels_vs_oracle(shape):
    r = T.rng(77)
    B, C, H, W = shape
    edge = (r.random((B, 1, H, W)) > 0.8).astype(np.float32)
    base = (r.random(shape, dtype=np.float32) * 1.2 - 0.1).astype(np.float32)
    g_out = r.standard_normal(shape, dtype=np.float32)
    for w in (1.0, 0.7):
        assert same(F_ee.gf_blend(cu(edge), cu(base), GAUSS, w), O.gf_blend_fwd(edge, base, w))
        g_edge, g_base = F_ee.gf_blend_backward(cu(g_out), cu(edge), cu(base), GAUSS, w)
        o_ge, o_gb = O.gf_blend_bwd(g_out, edge, base, w)
        assert same(g_edge, o_ge) and same(g_base, o_gb)
    only_base = F_ee.gf_blend_backward(cu(g_out), cu(edge), cu(base), GAUSS, 1.0, need_edge=False)
    assert only_base[0] is None and same(only_base[1], O.gf_blend_bwd(g_out, edge, base, 1.0)[1])


# ---------------------------------------------------------------------------------------------
# teacher-forced replay of the reference's PGD-10
# ---------------------------------------------------------------------------------------------
class TinyEENet(torch.nn.Module):
    """oracle/make_golden.py::TinyEENet with the drop-in filter (fused front end) and the same seeded head."""

    def __init__(self, canny, low, high, w, C, H, W, n_class, seed):
        super().__init__()
        self.canny, self.low, self.high, self.w = canny, low, high, w
        r = np.random.default_rng(seed)
        self.weight = torch.from_numpy(r.standard_normal((n_class, C * H * W)).astype(np.float32) * 0.05).to(DEV)

    def forward(self, x, want_edge=False):
        e = self.canny(x, low_threshold=self.low, high_threshold=self.high, hysteresis=True)
        z = torch.clamp(x + self.w * e, 0.0, 1.0)
        logits = z.reshape(z.shape[0], -1) @ self.weight.t()
        return (logits, e) if want_edge else logits


@pytest.mark.parametrize("variant", ["step125", "canny", "bpda"])
def test_teacher_forced_pgd10_against_reference_run(variant):
    """utils.attacks.PGD (reference, CPU fixture with every iterate's gradient and mask) replayed on the GPU drop-ins:
    mask exact and gradient within 1e-5 at every reference iterate, the fused step bit-exact given the reference's
    gradient, final iterate == the reference's x_adv bit for bit.  The free-running attacks.PGD is then compared with
    north_star's criterion (differences only where |g| < 1e-12 at some iteration) and the counts are printed."""
    z = np.load(os.path.join(GOLD, "pgd10_traced_%s.npz" % variant))
    x0, y, gs = cu(z["x"]), torch.from_numpy(z["y"]).to(DEV), z["gs"]
    B, C, H, W = x0.shape
    steps = gs.shape[0]
    edges = np.unpackbits(z["edges"])[:steps * B * H * W].reshape(steps, B, 1, H, W).astype(np.float32)
    torch.backends.cuda.matmul.allow_tf32 = False
    model = TinyEENet(quiet(CLS[variant], use_cuda=False, alpha=0.0), 38 / 255, 76 / 255, 1.0, C, H, W,
                      int(z["n_class"]), int(z["head_seed"]))
    x = x0.clone()
    worst = 0.0
    for i in range(steps):
        xi = x.detach().requires_grad_()
        logits, e = model(xi, want_edge=True)
        loss = torch.nn.functional.cross_entropy(logits, y, reduction='sum')
        g = torch.autograd.grad(loss, [xi])[0].cpu().numpy()
        assert np.array_equal(e.detach().cpu().numpy(), edges[i]), "iteration %d: mask differs" % i
        rel = np.abs(g - gs[i]).max() / np.abs(gs[i]).max()
        worst = max(worst, rel)
        assert rel <= 1e-5, (i, rel)
        x = F_ee.pgd_linf_step(x, cu(gs[i]), x0, 2 / 255, 16 / 255)          # teacher-forced
    assert np.array_equal(x.cpu().numpy(), z["x_adv"])

    class Args:
        random = False
        epsilon = 16 / 255
    grads = []

    def traced(xx):
        if xx.requires_grad:
            xx.register_hook(lambda g_: grads.append(g_.detach().clone()))
        return model(xx)
    free = attacks.PGD(traced, Args, x0, y, steps, 2 / 255).cpu().numpy()
    bad = free != z["x_adv"]
    ambiguous = np.zeros_like(bad)
    for ours, ref in zip(grads, gs):
        ambiguous |= (np.abs(ours.cpu().numpy()) < 1e-12) | (np.abs(ref) < 1e-12)
    # an element may also differ because an EARLIER ambiguous step moved a neighbour inside its 5x5 stencil: report both
    print("\n[%s] free-running PGD-10 vs reference: %d / %d elements differ, %d of them never sign-ambiguous themselves; "
          "worst teacher-forced gradient error %.2e of max|g|" % (variant, bad.sum(), bad.size, (bad & ~ambiguous).sum(), worst))
    assert np.abs(free - z["x"]).max() <= 16 / 255 + 1e-6
    assert bad.mean() < 0.02


# ---------------------------------------------------------------------------------------------
# PGD-L2 (cluster kernel) and the one-call iteration
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(8, 3, 64, 64), (5, 1, 28, 28), (3, 3, 32, 32), (3, 3, 224, 224), (2, 3, 128, 128),
                                   (2, 3, 288, 288), (4, 12), (2, 3, 33, 35), (300, 3, 64, 64)], ids=str)
def test_pgd_l2_cluster_step(shape):
    x, g, x0 = T.make_attack_inputs(19, shape, 0.05)
    for step, eps in ((0.5, 0.01), (0.003, 5.0)):
        got = F_ee.pgd_l2_step(cu(x), cu(g), cu(x0), step, eps)
        assert same(got, O.pgd_l2_step(x, g, x0, step, eps)), (shape, step)
    # generic three-pass path (forced): its own fixed order
    _lib.load().ee_set_tuning(0, 0, 1)
    assert same(F_ee.pgd_l2_step(cu(x), cu(g), cu(x0), 0.5, 0.01), O.pgd_l2_step(x, g, x0, 0.5, 0.01, three_pass=True))


def test_pgd_l2_repeatable_and_close_to_torch():
    shape = (64, 3, 64, 64)
    x, g, x0 = [cu(a) for a in T.make_attack_inputs(23, shape, 0.05)]
    a = F_ee.pgd_l2_step(x, g, x0, 0.5, 0.01)
    for _ in range(5):
        assert torch.equal(F_ee.pgd_l2_step(x, g, x0, 0.5, 0.01), a)
    gn = attacks.l2_norm(g).view(-1, 1, 1, 1) + 1e-8
    xa = x + 0.5 * (g / gn)
    d = xa - x0
    dn = attacks.l2_norm(d)
    cond = dn > 0.01
    d[cond] *= 0.01 / dn[cond].view(-1, 1, 1, 1)
    np.testing.assert_allclose(a.cpu().numpy(), torch.clamp(x0 + d, 0, 1).cpu().numpy(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("variant", ["step125", "canny", "bpda"])
@pytest.mark.parametrize("shape", [(4, 3, 64, 64), (3, 1, 28, 28), (2, 3, 224, 224), (2, 3, 17, 23)], ids=str)
def test_pgd_iteration_single_call(variant, shape):
    """ee_edge_pgd_iteration_f32 == the three separate entry points, bit for bit."""
    x, base, g_out, _ = T.make_inputs(41, *shape)
    x0 = np.clip(x + 0.02, 0, 1).astype(np.float32)
    low = None if variant == "step125" else T.LOW
    p = F_ee.make_params(variant, GAUSS, 0.0, low, T.HIGH, True)
    out, g_x, g_base, x_next = F_ee.pgd_iteration(cu(x), cu(base), cu(g_out), cu(x0), p, 1.0, 2 / 255, 16 / 255)
    assert torch.equal(out, F_ee.edge_blend(cu(x), cu(base), p, 1.0))
    r_gx, r_gb = F_ee.edge_blend_backward(cu(g_out), cu(x), cu(base), p, 1.0)
    assert torch.equal(g_x, r_gx) and torch.equal(g_base, r_gb)
    assert torch.equal(x_next, F_ee.pgd_linf_step(cu(x), r_gx, cu(x0), 2 / 255, 16 / 255))
    o2, _, gb2, xn2 = F_ee.pgd_iteration(cu(x), cu(base), cu(g_out), cu(x0), p, 1.0, 2 / 255, 16 / 255, want_out=False, want_base=False)
    assert o2 is None and gb2 is None and torch.equal(xn2, x_next)


def test_step125_params_are_normalised():
    """the fused module path used to send low / hysteresis to the step125 kernels."""
    f = quiet(core.CannyFilter_step125_1, use_cuda=False, alpha=0.0)
    p = f.params(0.1, 0.3, True)
    assert p.has_low == 0 and p.hysteresis == 0 and p.has_high == 1
    x, base, g_out, _ = T.make_inputs(5, 2, 3, 64, 64)
    a = F_ee.edge_blend(cu(x), cu(base), p, 1.0)
    q = F_ee.make_params("step125", GAUSS, 0.0, 0.1, 0.3, True)          # raw struct with the fields set: same result
    assert torch.equal(a, F_ee.edge_blend(cu(x), cu(base), q, 1.0))


def test_hfs_module_has_no_silent_fallback():
    m = core.HighFreqSuppress(96, 96, 12)
    with pytest.raises(RuntimeError):
        m(torch.rand((2, 3, 96, 96), device=DEV))
    with pytest.raises(RuntimeError):
        core.HighFreqSuppress(64, 64, 8)(torch.rand(2, 3, 64, 64))              # CPU tensor
    x = torch.rand((2, 3, 96, 96), device=DEV)
    y = core.HighFreqSuppress(96, 96, 12, impl='torch_fft')(x)                   # the explicit opt-in
    assert y.shape == x.shape
    with pytest.raises(NotImplementedError):
        core.HighFreqSuppress(64, 64, 8, c2r='full')


# ---------------------------------------------------------------------------------------------
# strided tensors at the C ABI (ee_*_strided_f32): views are read in place, results identical to the dense call
# ---------------------------------------------------------------------------------------------
def _views(B, C, H, W, seed):
    """name -> (x, base, g_out) views of larger tensors, all with unit column stride but not contiguous."""
    gen = torch.Generator(device=DEV).manual_seed(seed)
    def big(*shape):
        return torch.rand(shape, device=DEV, generator=gen)
    out = {}
    t = [big(2 * B, C, H, W) for _ in range(3)]
    out["batch_step2"] = tuple(v[::2] for v in t)
    t = [big(B, C + 2, H, W) for _ in range(3)]
    out["channel_slice"] = tuple(v[:, 1:1 + C] for v in t)
    t = [big(B, C, H + 8, W + 8) for _ in range(3)]
    out["crop_aligned"] = tuple(v[:, :, 4:4 + H, 4:4 + W] for v in t)          # 16-byte aligned rows: 128-bit path
    out["crop_unaligned"] = tuple(v[:, :, 3:3 + H, 5:5 + W] for v in t)        # scalar path
    t = [big(1, C, H, W) for _ in range(3)]
    out["expanded_batch"] = tuple(v.expand(B, C, H, W) for v in t)
    return out


@pytest.mark.parametrize("variant", ["step125", "canny", "bpda"])
@pytest.mark.parametrize("shape", [(4, 3, 64, 64), (3, 1, 28, 28), (2, 3, 24, 40)], ids=str)
def test_strided_views_are_read_in_place(variant, shape):
    B, C, H, W = shape
    low = None if variant == "step125" else T.LOW
    p = F_ee.make_params(variant, GAUSS, 0.0, low, T.HIGH, True)
    for name, (x, base, g_out) in _views(B, C, H, W, 3).items():
        assert not x.is_contiguous(), name
        base = base * 1.1 - 0.1
        g_out = g_out - 0.5
        xc, bc, gc = x.contiguous(), base.contiguous(), g_out.contiguous()
        want_out, want_edge = F_ee.edge_blend(xc, bc, p, 1.0, want_edge=True)
        got_out, got_edge = F_ee.edge_blend(x, base, p, 1.0, want_edge=True)
        assert torch.equal(got_out, want_out) and torch.equal(got_edge, want_edge), name
        want_gx, want_gb = F_ee.edge_blend_backward(gc, xc, bc, p, 1.0)
        got_gx, got_gb = F_ee.edge_blend_backward(g_out, x, base, p, 1.0)
        assert torch.equal(got_gx, want_gx) and torch.equal(got_gb, want_gb), name
        assert torch.equal(F_ee.edge_map(x, p), F_ee.edge_map(xc, p)), name
        ge = g_out[:, :1]
        assert torch.equal(F_ee.edge_map_backward(ge, x, p), F_ee.edge_map_backward(ge.contiguous(), xc, p)), name
    # strided OUTPUT buffers
    x, base, g_out = (v.contiguous() for v in _views(B, C, H, W, 4)["batch_step2"])
    big_out = torch.zeros((B, C, H + 4, W + 4), device=DEV)
    view = big_out[:, :, 2:2 + H, 0:W]
    F_ee.edge_blend(x, base[::1], p, 1.0, out=None)
    got = F_ee._edge_blend_strided(x, base, p, 1.0, False, view)
    assert got is view and torch.equal(view, F_ee.edge_blend(x, base, p, 1.0)) and float(big_out[:, :, :2].abs().sum()) == 0.0


def test_strided_autograd_and_unsupported_strides():
    f = quiet(core.CannyFilter_step125_1, use_cuda=False, alpha=0.0)
    big = torch.rand((8, 3, 64, 64), device=DEV)
    xv = big[1::2].detach().requires_grad_()            # a non-contiguous leaf view
    xc = big[1::2].contiguous().requires_grad_()
    g = torch.randn((4, 3, 64, 64), device=DEV)
    core.edge_enhance(xv, xv, f, 1.0, None, T.HIGH, True).backward(g)
    core.edge_enhance(xc, xc, f, 1.0, None, T.HIGH, True).backward(g)
    assert torch.equal(xv.grad, xc.grad)
    # a transposed view (column stride != 1) is copied by the wrapper, and rejected at the ABI
    xt = torch.rand((2, 3, 64, 64), device=DEV).transpose(2, 3)
    p = f.params(None, T.HIGH, False)
    assert torch.equal(F_ee.edge_map(xt, p), F_ee.edge_map(xt.contiguous(), p))
    L = _lib.load()
    import ctypes
    s = _lib.EEStrides(*xt.stride())
    edge = torch.empty((2, 1, 64, 64), device=DEV)
    rc = L.ee_edge_fwd_strided_f32(xt.data_ptr(), ctypes.byref(s), edge.data_ptr(), None, 2, 3, 64, 64, ctypes.byref(p), None)
    assert rc == -2 and b"column stride" in L.ee_last_error()


# ---------------------------------------------------------------------------------------------
# edge cases of the round-2 entry points: empty batches, degenerate shapes, loud errors
# ---------------------------------------------------------------------------------------------
def test_round2_entry_points_edge_cases():
    p = F_ee.make_params("step125", GAUSS, 0.0, None, T.HIGH, False)
    empty = torch.empty((0, 3, 8, 8), device=DEV)
    assert F_ee.pgd_l2_step(empty, empty, empty, 0.1, 0.1).shape == (0, 3, 8, 8)
    assert F_ee.gf_blend(torch.empty((0, 1, 8, 8), device=DEV), empty, GAUSS, 1.0).shape == (0, 3, 8, 8)
    out, g_x, g_base, x_next = F_ee.pgd_iteration(empty, empty, empty, empty, p, 1.0, 0.01, 0.1)
    assert x_next.shape == (0, 3, 8, 8)
    # degenerate planes through the gf kernels (1 x 1, one row, one column)
    r = T.rng(5)
    for shape in ((1, 1, 1, 1), (2, 3, 1, 8), (2, 2, 9, 1)):
        B, C, H, W = shape
        edge = (r.random((B, 1, H, W)) > 0.5).astype(np.float32)
        base = r.random(shape, dtype=np.float32)
        g = r.standard_normal(shape, dtype=np.float32)
        assert same(F_ee.gf_blend(cu(edge), cu(base), GAUSS, 1.0), O.gf_blend_fwd(edge, base, 1.0))
        ge, gb = F_ee.gf_blend_backward(cu(g), cu(edge), cu(base), GAUSS, 1.0)
        oge, ogb = O.gf_blend_bwd(g, edge, base, 1.0)
        assert same(ge, oge) and same(gb, ogb)
    # one-pass PGD-L2 on a single sample and on a sample of 4 elements
    for shape in ((1, 3, 64, 64), (3, 4), (1, 3, 224, 224)):
        x, g, x0 = T.make_attack_inputs(2, shape, 0.05)
        assert same(F_ee.pgd_l2_step(cu(x), cu(g), cu(x0), 0.5, 0.01), O.pgd_l2_step(x, g, x0, 0.5, 0.01))
    # loud errors
    x = torch.rand((2, 3, 8, 8), device=DEV)
    with pytest.raises(RuntimeError):
        F_ee.pgd_l2_step(x.cpu(), x.cpu(), x.cpu(), 0.1, 0.1)
    with pytest.raises(ValueError):
        F_ee.gf_blend(torch.rand((2, 1, 8, 9), device=DEV), x, GAUSS, 1.0)
    bad = GAUSS.copy(); bad[0, 0] += 1
    with pytest.raises(RuntimeError):
        F_ee.gf_blend(torch.rand((2, 1, 8, 8), device=DEV), x, bad, 1.0)
    q = F_ee.make_params("step125", GAUSS, 0.0, None, T.HIGH, False)
    q.flags = 4
    with pytest.raises(RuntimeError):
        F_ee.edge_map(x, q)
