"""-m gpu: CUDA kernels (through the C ABI) vs the CPU oracle on the same seeded inputs.

The oracle and the kernels evaluate the same canonical fp32 expression trees, so everything is
compared BIT-EXACT: edge masks, blended images, gradients and attack updates."""
import numpy as np
import pytest
import torch

from tests import common as T

pytestmark = pytest.mark.gpu

ee = pytest.importorskip("edge_enhancement_b200")
from edge_enhancement_b200 import functional as F_ee   # noqa: E402
from edge_enhancement_b200 import _lib                 # noqa: E402
from oracle import oracle as O                         # noqa: E402

DEV = "cuda:0"
GAUSS = O.gaussian3(0.0, 1.0)


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def both_params(variant, alpha, low, high, hyst, sigma=1.0):
    g = O.gaussian3(0.0, sigma)
    return (F_ee.make_params(variant, g, alpha, low, high, hyst),
            O.make_params(variant, sigma=sigma, alpha=alpha, low=low, high=high, hysteresis=hyst))


@pytest.fixture(autouse=True)
def _reset_tuning():
    _lib.load().ee_set_tuning(0, 0, 0)
    yield
    _lib.load().ee_set_tuning(0, 0, 0)


def assert_same(name, got, want):
    got = got.cpu().numpy() if isinstance(got, torch.Tensor) else got
    if got.shape != want.shape:
        raise AssertionError("%s: shape %s vs %s" % (name, got.shape, want.shape))
    same = (got == want) | (np.isnan(got) & np.isnan(want))
    if not same.all():
        bad = np.argwhere(~same)
        i = tuple(bad[0])
        raise AssertionError("%s: %d / %d elements differ; first at %s: cuda=%r oracle=%r; max abs diff %g"
                             % (name, len(bad), got.size, i, got[i], want[i],
                                np.nanmax(np.abs(got.astype(np.float64) - want.astype(np.float64)))))


SHAPES = [
    # B, C, H, W      (W % 4 == 0 -> 128-bit path, otherwise scalar path)
    (2, 3, 64, 64), (3, 1, 28, 28), (2, 3, 32, 32), (1, 3, 224, 224), (2, 3, 17, 23), (2, 2, 9, 12),
    (3, 1, 1, 1), (2, 3, 1, 8), (2, 3, 8, 1), (1, 4, 2, 2), (1, 3, 5, 4), (1, 3, 40, 300),
    (2, 3, 4, 8), (1, 5, 13, 16), (1, 3, 9, 1028), (2, 3, 12, 132), (1, 1, 30, 256),
    (1, 1, 64, 256), (1, 2, 32, 200), (1, 3, 288, 288),        # wide + H % 4 == 0: chunk-aligned tile backward
]
VARIANT_MODES = [("step125", "hyst")] + [(v, m) for v in ("canny", "bpda") for m in ("hyst", "mix", "low", "raw")]


# staging knob of ee_set_tuning: 0 = auto, 1 = generic kernels only, 3 = strip kernels instead of the chunk-aligned
# tiles for wide images, 4 = tuned kernels also for wide images, 6 / 7 = always / never stage x tiles by TMA tensor copies
@pytest.mark.parametrize("staging", [0, 1, 3, 4, 6, 7, 8, 9])
@pytest.mark.parametrize("strip", [0, 1, 3, 7])
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("variant,mode", VARIANT_MODES)
def test_edge_filter_fwd_bwd(variant, mode, shape, strip, staging):
    B, C, H, W = shape
    low, high, hyst = T.MODES[mode]
    alpha = 0.05 if variant != "bpda" else 0.0
    pc, po = both_params(variant, alpha, low, high, hyst)
    x, base, g_out, g_edge = T.make_inputs(hash((variant, mode, shape)) % 10000, B, C, H, W)
    _lib.load().ee_set_tuning(strip, strip, staging)
    # module-level forward / backward
    assert_same("edge", F_ee.edge_map(cu(x), pc), O.edge_fwd(x, po))
    assert_same("g_x(edge)", F_ee.edge_map_backward(cu(g_edge), cu(x), pc), O.edge_bwd(g_edge, x, po))
    # fused with the blend
    w = 1.0
    out, edge = F_ee.edge_blend(cu(x), cu(base), pc, w, want_edge=True)
    o_out, o_edge = O.edge_blend_fwd(x, base, po, w, want_edge=True)
    assert_same("blend edge", edge, o_edge)
    assert_same("blend out", out, o_out)
    g_x, g_base = F_ee.edge_blend_backward(cu(g_out), cu(x), cu(base), pc, w)
    o_gx, o_gb = O.edge_blend_bwd(g_out, x, base, po, w)
    assert_same("g_base", g_base, o_gb)
    assert_same("g_x(blend)", g_x, o_gx)


@pytest.mark.parametrize("kind", ["sparse", "smooth"])
@pytest.mark.parametrize("variant", ["step125", "canny", "bpda"])
def test_edge_filter_other_inputs(variant, kind):
    # MNIST config: alpha 0.3, low/high 25/51 (configs_mnist/ee_at_training.yml) on mostly-flat images
    B, C, H, W = (4, 1, 28, 28) if kind == "sparse" else (2, 3, 64, 64)
    alpha = 0.3 if variant != "bpda" else 0.0
    pc, po = both_params(variant, alpha, 25 / 255, 51 / 255, True)
    x, base, g_out, _ = T.make_inputs(7, B, C, H, W, kind)
    for w in (1.0, 0.5):
        assert_same("out", F_ee.edge_blend(cu(x), cu(base), pc, w), O.edge_blend_fwd(x, base, po, w))
        g_x, g_base = F_ee.edge_blend_backward(cu(g_out), cu(x), cu(base), pc, w)
        o_gx, o_gb = O.edge_blend_bwd(g_out, x, base, po, w)
        assert np.isfinite(o_gx).all()              # sub-gradient at mag == 0 is 0, never NaN
        assert_same("g_base", g_base, o_gb)
        assert_same("g_x", g_x, o_gx)


@pytest.mark.parametrize("sigma", [0.5, 2.0])
def test_other_sigma_and_negative_threshold(sigma):
    x, base, g_out, g_edge = T.make_inputs(11, 2, 3, 32, 32)
    for variant in ("step125", "bpda", "canny"):
        pc, po = both_params(variant, 0.0, -0.2 if variant != "step125" else None, -0.1, True, sigma=sigma)
        assert_same("edge", F_ee.edge_map(cu(x), pc), O.edge_fwd(x, po))
        assert_same("g_x", F_ee.edge_map_backward(cu(g_edge), cu(x), pc), O.edge_bwd(g_edge, x, po))


def test_unaligned_views_take_scalar_path():
    # a 16-byte-misaligned (but contiguous) tensor must still give the same answer
    x, base, g_out, _ = T.make_inputs(5, 2, 3, 16, 16)
    pc, po = both_params("step125", 0.0, None, T.HIGH, False)
    n = x.size
    buf = torch.empty(n + 1, device=DEV)
    xs = buf[1:].view(2, 3, 16, 16)
    xs.copy_(cu(x))
    assert xs.data_ptr() % 16 != 0 and xs.is_contiguous()
    assert_same("out", F_ee.edge_blend(xs, cu(base), pc, 1.0), O.edge_blend_fwd(x, base, po, 1.0))


def test_only_g_base_requested():
    x, base, g_out, _ = T.make_inputs(6, 2, 3, 32, 32)
    for variant in ("step125", "canny"):
        pc, po = both_params(variant, 0.0, T.LOW, T.HIGH, True)
        g_x, g_base = F_ee.edge_blend_backward(cu(g_out), cu(x), cu(base), pc, 1.0, need_x=False, need_base=True)
        assert g_x is None
        assert_same("g_base", g_base, O.edge_blend_bwd(g_out, x, base, po, 1.0)[1])
        g_x, g_base = F_ee.edge_blend_backward(cu(g_out), cu(x), cu(base), pc, 1.0, need_x=True, need_base=False)
        assert g_base is None
        assert_same("g_x", g_x, O.edge_blend_bwd(g_out, x, base, po, 1.0)[0])


def test_cluster_backward_224():
    """ImageNet size: the default chunk-aligned tile backward (ee_edge_tiles.cuh), the older strip kernels (staging 3)
    and the opt-in thread-block-cluster backward (8 CTAs x 28 rows, halo rows through distributed shared memory,
    ee_edge_cluster.cuh; staging 5) and the TMA-staged tiles (staging 6) all give the oracle's bits, also when only one of the two gradients is requested."""
    x, base, g_out, g_edge = T.make_inputs(224, 3, 3, 224, 224)
    pc, po = both_params("step125", 0.02, None, T.HIGH, False)
    L = _lib.load()
    o_gx, o_gb = O.edge_blend_bwd(g_out, x, base, po, 1.0)
    for staging in (0, 3, 5, 6, 7):
        L.ee_set_tuning(0, 0, staging)
        g_x, g_base = F_ee.edge_blend_backward(cu(g_out), cu(x), cu(base), pc, 1.0)
        assert_same("g_base", g_base, o_gb)
        assert_same("g_x", g_x, o_gx)
        only_b = F_ee.edge_blend_backward(cu(g_out), cu(x), cu(base), pc, 1.0, need_x=False, need_base=True)
        assert only_b[0] is None
        assert_same("g_base only", only_b[1], o_gb)
        only_x = F_ee.edge_blend_backward(cu(g_out), cu(x), cu(base), pc, 1.0, need_x=True, need_base=False)
        assert only_x[1] is None
        assert_same("g_x only", only_x[0], o_gx)
        assert_same("g_x(edge)", F_ee.edge_map_backward(cu(g_edge), cu(x), pc), O.edge_bwd(g_edge, x, po))
    L.ee_set_tuning(0, 0, 0)


# ---------------------------------------------------------------------------------------------
# attack updates: bit-exact vs the oracle AND vs the torch expression of the reference on the GPU
# ---------------------------------------------------------------------------------------------
ATTACK_SHAPES = [(256, 3, 64, 64), (5, 3, 7, 9), (1,), (3,), (4099,)]


@pytest.mark.parametrize("shape", ATTACK_SHAPES, ids=str)
def test_pgd_linf_step(shape):
    eps, a = 16 / 255, 2 / 255
    x, g, x0 = T.make_attack_inputs(3, shape, eps)
    if g.size > 8:
        g.reshape(-1)[5] = np.nan
    for alpha in (a, -a):
        want = O.pgd_linf_step(x, g, x0, alpha, eps)
        got = F_ee.pgd_linf_step(cu(x), cu(g), cu(x0), alpha, eps)
        assert_same("pgd", got, want)
        tx, tg, tx0 = cu(x), cu(g), cu(x0)
        ref = tx + alpha * torch.sign(tg)
        ref = torch.min(torch.max(ref, tx0 - eps), tx0 + eps)
        ref = torch.clamp(ref, 0, 1)                      # utils/attacks.py:25-27 verbatim
        assert_same("pgd vs torch", got, ref.cpu().numpy())
        assert float(got.min()) >= 0 and float(got.max()) <= 1
    # in place
    tx = cu(x)
    F_ee.pgd_linf_step(tx, cu(g), cu(x0), a, eps, out=tx)
    assert_same("pgd in-place", tx, O.pgd_linf_step(x, g, x0, a, eps))


@pytest.mark.parametrize("shape", ATTACK_SHAPES, ids=str)
def test_fgsm_free_cw_steps(shape):
    eps, a = 4 / 255, 4 / 255
    x, g, x0 = T.make_attack_inputs(4, shape, eps)
    assert_same("fgsm", F_ee.fgsm_step(cu(x), cu(g), -a), O.fgsm_step(x, g, -a))
    delta = (x - x0).astype(np.float32)
    od, oadv = O.free_at_step(delta, g, x0, a, eps)
    td = cu(delta)
    adv = F_ee.free_at_step_(td, cu(g), cu(x0), a, eps)
    assert_same("free delta", td, od)
    assert_same("free adv", adv, oadv)
    tdd = cu(delta); tdd += a * torch.sign(cu(g)); tdd.clamp_(-eps, eps)
    assert_same("free delta vs torch", td, tdd.cpu().numpy())
    mn, mx = (x0 - 0.02).astype(np.float32), (x0 + 0.02).astype(np.float32)
    assert_same("cw", F_ee.cw_linf_step(cu(x), cu(g), cu(x0), cu(mn), cu(mx), 0.00392, 0.03),
                O.cw_linf_step(x, g, x0, mn, mx, 0.00392, 0.03))


@pytest.mark.parametrize("shape", [(4, 3, 64, 64), (3, 1, 28, 28), (2, 5), (2, 3, 33, 35)], ids=str)
def test_pgd_l2_step(shape):
    x, g, x0 = T.make_attack_inputs(9, shape, 0.05)
    for step, eps in ((0.5, 0.01), (0.003, 5.0)):
        assert_same("l2", F_ee.pgd_l2_step(cu(x), cu(g), cu(x0), step, eps), O.pgd_l2_step(x, g, x0, step, eps))


def test_ste_functions():
    r = T.rng(1)
    v = (r.standard_normal(5000, dtype=np.float32) * 0.8).astype(np.float32)
    v[:10] = 0.5; v[10:20] = 0.0; v[20] = 1.001; v[21] = np.float32(1.0010001)
    g = r.standard_normal(5000, dtype=np.float32)
    from edge_enhancement_b200 import core
    for thr in (0.3, -0.3):
        t = cu(v).requires_grad_()
        out = core.To_compare.apply(t, torch.tensor(thr))
        out.backward(cu(g))
        assert_same("cmp fwd", out.detach(), O.to_compare_fwd(v, thr))
        assert_same("cmp bwd", t.grad, O.to_compare_bwd(g, v, thr))
    t = cu(v).requires_grad_(); out = core.To_eq.apply(t); out.backward(cu(g))
    assert_same("eq fwd", out.detach(), O.to_eq_fwd(v)); assert_same("eq bwd", t.grad, O.to_eq_bwd(g, v))
    t = cu(v).requires_grad_(); out = core.BinaryConnectDeterministic.apply(t); out.backward(cu(g))
    assert_same("sign fwd", out.detach(), O.safe_sign_fwd(v)); assert_same("sign bwd", t.grad, O.safe_sign_bwd(g, v))
    assert_same("safeSign", core.safeSign(cu(v)), O.safe_sign_fwd(v))


def test_errors_are_loud():
    p = F_ee.make_params("step125", GAUSS, 0.0, None, None, False)     # no high threshold
    with pytest.raises(RuntimeError):
        F_ee.edge_map(torch.rand(1, 3, 8, 8, device=DEV), p)
    with pytest.raises(RuntimeError):
        F_ee.edge_map(torch.rand(1, 3, 8, 8), F_ee.make_params("canny", GAUSS))          # CPU tensor
    with pytest.raises(TypeError):
        F_ee.edge_map(torch.rand(1, 3, 8, 8, device=DEV).double(), F_ee.make_params("canny", GAUSS))
    bad = GAUSS.copy(); bad[0, 0] += 1
    with pytest.raises(RuntimeError):
        F_ee.edge_map(torch.rand(1, 3, 8, 8, device=DEV), F_ee.make_params("canny", bad))


@pytest.mark.parametrize("variant", ["step125", "canny", "bpda"])
@pytest.mark.parametrize("shape", [(3, 3, 64, 64), (2, 3, 24, 40), (1, 3, 40, 300), (2, 3, 7, 10)], ids=str)
def test_channels_last_inputs(variant, shape):
    """torch.channels_last (NHWC) tensors go through the fused kernels without a layout copy (C == 3,
    W % 4 == 0) or through an NCHW copy otherwise; results equal the NCHW path bit for bit."""
    B, C, H, W = shape
    low = None if variant == "step125" else T.LOW
    pc, po = both_params(variant, 0.0, low, T.HIGH, True)
    x, base, g_out, _ = T.make_inputs(21, B, C, H, W)
    cl = lambda a: cu(a).contiguous(memory_format=torch.channels_last)
    out = F_ee.edge_blend(cl(x), cl(base), pc, 1.0)
    g_x, g_base = F_ee.edge_blend_backward(cl(g_out), cl(x), cl(base), pc, 1.0)
    if W % 4 == 0:
        assert out.is_contiguous(memory_format=torch.channels_last) and g_x.is_contiguous(memory_format=torch.channels_last)
    o_out = O.edge_blend_fwd(x, base, po, 1.0)
    o_gx, o_gb = O.edge_blend_bwd(g_out, x, base, po, 1.0)
    assert_same("out", out.contiguous(), o_out)
    assert_same("g_base", g_base.contiguous(), o_gb)
    assert_same("g_x", g_x.contiguous(), o_gx)
    # and through autograd with a channels_last model input
    from edge_enhancement_b200 import core
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        f = {"step125": core.CannyFilter_step125_1, "canny": core.CannyFilter, "bpda": core.CannyFilter_BPDA}[variant]()
    xt = cl(x).requires_grad_()
    bt = cl(base).requires_grad_()
    y = core.edge_enhance(xt, bt, f, 1.0, low, T.HIGH, True)
    y.backward(cl(g_out))
    assert_same("autograd g_x", xt.grad.contiguous(), o_gx)
    assert_same("autograd g_base", bt.grad.contiguous(), o_gb)
