"""not gpu: pin the C oracle against the golden fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py ran utils/core.py + utils/attacks.py on CPU in the build container).

Tolerances (BASELINE.json north_star): threshold masks exact; edge maps / blended images 1e-5
relative; gradients 1e-5 relative to the tensor's max |g| on positions where the reference gradient
is finite (the reference produces NaN where mag == 0; the oracle defines 0)."""
import glob
import os

import numpy as np
import pytest

from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EDGE_FILES = sorted(glob.glob(os.path.join(GOLD, "edge_*.npz")))


def _opt(s):
    return None if s == "None" else float(s)


def load_edge_case(path):
    z = np.load(path)
    variant, alpha, sigma, low, high, hyst, w = [str(v) for v in z["meta"]]
    params = dict(variant=variant, alpha=float(alpha), sigma=float(sigma), low=_opt(low), high=_opt(high),
                  hysteresis=bool(int(hyst)))
    return z, params, float(w)


def check_against_reference(name, edge, out, g_x, g_base, z, binary):
    if binary:
        assert np.array_equal(edge, z["edge"]), "%s: %d mask pixels differ" % (name, (edge != z["edge"]).sum())
    else:
        np.testing.assert_allclose(edge, z["edge"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(out, z["out"], rtol=1e-5, atol=1e-7)
    assert np.array_equal(g_base, z["g_base"])        # a mask applied to g_out: exact
    ref = z["g_x"]
    fin = np.isfinite(ref)
    assert fin.mean() > 0.5
    scale = np.abs(ref[fin]).max() if fin.any() else 1.0
    assert np.isfinite(g_x).all()
    assert np.abs(g_x - ref)[fin].max() <= 1e-5 * max(scale, 1e-30), \
        "%s: g_x differs by %g (scale %g)" % (name, np.abs(g_x - ref)[fin].max(), scale)


def test_fixtures_present():
    assert len(EDGE_FILES) >= 18
    assert os.path.exists(os.path.join(GOLD, "attack_steps.npz"))


@pytest.mark.parametrize("path", EDGE_FILES, ids=lambda p: os.path.basename(p)[5:-4])
def test_oracle_matches_reference_fixture(path):
    z, kw, w = load_edge_case(path)
    p = O.make_params(**kw)
    out, edge = O.edge_blend_fwd(z["x"], z["base"], p, w, want_edge=True)
    g_x, g_base = O.edge_blend_bwd(z["g_out"], z["x"], z["base"], p, w)
    binary = not (kw["low"] is None or (kw["variant"] == "bpda" and kw["high"] is None))
    check_against_reference(os.path.basename(path), edge, out, g_x, g_base, z, binary)
    # module-level forward agrees with the fused one
    assert np.array_equal(O.edge_fwd(z["x"], p), edge)


def test_attack_steps_bit_exact():
    z = np.load(os.path.join(GOLD, "attack_steps.npz"))
    eps, a = 16 / 255, 2 / 255
    assert np.array_equal(O.pgd_linf_step(z["x"], z["g"], z["x0"], a, eps), z["pgd"])
    assert np.array_equal(O.pgd_linf_step(z["x"], z["g"], z["x0"], -a, eps), z["tpgd"])
    assert np.array_equal(O.fgsm_step(z["x"], z["g"], 0.007), z["fgsm"])
    d, adv = O.free_at_step(z["x"] - z["x0"], z["g"], z["x0"], 4 / 255, 4 / 255)
    assert np.array_equal(d, z["free_delta"]) and np.array_equal(adv, z["free_adv"])
    assert np.array_equal(O.cw_linf_step(z["x"], z["g"], z["x0"], z["cw_min"], z["cw_max"], 0.00392, 0.02), z["cw"])
    # L2 step: per-sample mean reduction order differs from torch's -> tolerance, not bits
    np.testing.assert_allclose(O.pgd_l2_step(z["x"], z["g2"], z["x0"], 0.5, 0.02), z["l2"], rtol=1e-5, atol=1e-6)


def test_known_answers():
    """Hand-derivable cases (SURVEY.md section 4 'known-answer')."""
    p = O.make_params("step125", high=0.1)
    # constant image: no gradient anywhere -> no edges, including the replicate-padded border
    x = np.full((1, 3, 9, 9), 0.7, np.float32)
    assert O.edge_fwd(x, p).sum() == 0
    # vertical step edge: edges exactly on the two columns next to the step, all rows
    x = np.zeros((1, 1, 8, 12), np.float32); x[..., 6:] = 1.0
    e = O.edge_fwd(x, p)[0, 0]
    cols = np.where(e.any(axis=0))[0]
    assert set(cols) <= {3, 4, 5, 6, 7, 8} and {5, 6} <= set(cols)
    assert (e == e[0:1]).all()                     # identical in every row (replicate border)
    # Gaussian taps: fp32 constants of SURVEY.md appendix A.1
    g = O.gaussian3()
    assert g[0, 0] == np.float32(0.07511361) and g[0, 1] == np.float32(0.123841405) and g[1, 1] == np.float32(0.20417996)
    # gradient w.r.t. every channel is the same plane
    r = np.random.default_rng(0)
    x = r.random((2, 3, 12, 12), dtype=np.float32)
    gx = O.edge_bwd(r.standard_normal((2, 1, 12, 12), dtype=np.float32), x, p)
    assert np.array_equal(gx[:, 0], gx[:, 1]) and np.array_equal(gx[:, 0], gx[:, 2])


def test_adjoint_identity():
    """<J v, u> == <v, J^T u> for the linear part (blur + Sobel with replicate pads): checked through
    the STEP125 backward with a threshold that passes everything and mag-independent cotangents is
    not linear, so use finite differences on the smooth quantity sum(g_edge * mag) instead."""
    r = np.random.default_rng(3)
    x = r.random((1, 2, 7, 6)).astype(np.float32)
    ge = r.standard_normal((1, 1, 7, 6)).astype(np.float32)
    p = O.make_params("canny", low=None, high=None)          # raw thinned magnitude output
    # finite differences in float64 on a float32 pipeline are noisy; use a large step and loose bound
    gx = O.edge_bwd(ge, x, p)
    f0 = (O.edge_fwd(x, p) * ge).sum(dtype=np.float64)
    num = np.zeros_like(x, dtype=np.float64)
    h = 1e-2
    for idx in np.ndindex(*x.shape):
        xp = x.copy(); xp[idx] += h
        xm = x.copy(); xm[idx] -= h
        num[idx] = ((O.edge_fwd(xp, p) * ge).sum(dtype=np.float64) - (O.edge_fwd(xm, p) * ge).sum(dtype=np.float64)) / (2 * h)
    # NMS decisions can flip under the perturbation; compare where the analytic and numeric agree in bulk
    close = np.isclose(num, gx, rtol=0.05, atol=0.02)
    assert close.mean() > 0.7, close.mean()
    assert np.isfinite(f0)


SQUARE_FILES = sorted(glob.glob(os.path.join(GOLD, "add_square_*.npz")))


@pytest.mark.parametrize("path", SQUARE_FILES, ids=lambda p: os.path.basename(p)[11:-4])
def test_add_square_oracle_matches_reference_fixture(path):
    """Add_Square (utils/core.py:640-655): forward bit-exact; the gradient is g times a 0/1 multiplier -- exact for
    n_queries = 1 (every reference config), within 1 ulp of autograd's accumulation order otherwise."""
    z = np.load(path)
    eps, nq = float(str(z["meta"][2])), int(str(z["meta"][3]))
    assert len(SQUARE_FILES) >= 3
    out = O.add_square(z["x"], z["stripe"], z["table"], eps)
    assert np.array_equal(out, z["out"])
    g_x = O.add_square(z["x"], z["stripe"], z["table"], eps, g=z["g"])
    if nq == 1:
        assert np.array_equal(g_x, z["g_x"])
    else:
        np.testing.assert_allclose(g_x, z["g_x"], rtol=2e-7, atol=0)
    # properties: inside the eps ball and [0, 1]
    assert (np.abs(out - z["x"]) <= np.float32(eps) + 1e-7).all() and out.min() >= 0.0 and out.max() <= 1.0


def test_attack_extras_bit_exact():
    """random start (attacks.py:15-17) and AVmixup vertex / float64 mix (attacks.py:469-478)"""
    z = np.load(os.path.join(GOLD, "attack_extras.npz"))
    assert np.array_equal(O.add_clamp(z["x0"], z["noise"]), z["start"])
    assert np.array_equal(O.avmixup_mix(z["x"], z["x0"], z["weight"], float(z["gamma"])), z["mixed"])


# ---------------------------------------------------------------------------------------------
# round 2: NaN-compatible backward, standalone STE Functions, with_gf blend, teacher-forced PGD-10
# ---------------------------------------------------------------------------------------------
NAN_FILES = sorted(glob.glob(os.path.join(GOLD, "nan_*.npz")))


def check_nan_case(z, g_x, g_base, edge):
    """The reference's NaN set exactly; finite entries within 1e-5 of max |g|; mask and g_base exact."""
    ref = z["g_x"]
    assert np.array_equal(edge, z["edge"])
    assert np.array_equal(g_base, z["g_base"])
    assert np.isnan(ref).mean() > 0.05, "fixture has no flat region"
    assert np.array_equal(np.isnan(g_x), np.isnan(ref)), "%d entries differ in NaN-ness" % (np.isnan(g_x) != np.isnan(ref)).sum()
    fin = np.isfinite(ref)
    assert np.abs(g_x - ref)[fin].max() <= 1e-5 * np.abs(ref[fin]).max()


@pytest.mark.parametrize("path", NAN_FILES, ids=lambda p: os.path.basename(p)[4:-4])
def test_oracle_nan_compat_matches_reference_backward(path):
    """Flat / saturated regions: the reference's autograd returns NaN on the 5x5 neighbourhood of every mag == 0 pixel
    (ADVICE.md round 1).  nan_compat=True reproduces that set exactly; the default keeps the sub-gradient 0."""
    assert len(NAN_FILES) >= 4
    z, kw, w = load_edge_case(path)
    p = O.make_params(nan_compat=True, **kw)
    out, edge = O.edge_blend_fwd(z["x"], z["base"], p, w, want_edge=True)
    g_x, g_base = O.edge_blend_bwd(z["g_out"], z["x"], z["base"], p, w)
    check_nan_case(z, g_x, g_base, edge)
    # default mode: finite everywhere, equal to the NaN-compatible result wherever that is finite
    g0, _ = O.edge_blend_bwd(z["g_out"], z["x"], z["base"], O.make_params(**kw), w)
    assert np.isfinite(g0).all()
    fin = np.isfinite(g_x)
    assert np.array_equal(g0[fin], g_x[fin])


def test_ste_functions_oracle_matches_reference_fixture():
    """To_compare / To_eq / BinaryConnectDeterministic / safeSign stand-alone (utils/core.py:115-145, :329-382), incl. the
    negative-threshold quirk of To_compare.forward (two sequential masked writes: everything becomes 1)."""
    z = np.load(os.path.join(GOLD, "ste_functions.npz"))
    v, g = z["v"], z["g"]
    for tag in ("pos", "neg", "zero"):
        thr = float(z["cmp_%s_thr" % tag])
        assert np.array_equal(O.to_compare_fwd(v, thr), z["cmp_%s_fwd" % tag]), tag
        assert np.array_equal(O.to_compare_bwd(g, v, thr), z["cmp_%s_bwd" % tag]), tag
    assert z["cmp_neg_fwd"].min() == 1.0                   # the quirk is really in the fixture
    assert np.array_equal(O.to_eq_fwd(v), z["eq_fwd"]) and np.array_equal(O.to_eq_bwd(g, v), z["eq_bwd"])
    assert z["eq_fwd"].sum() >= 1
    assert np.array_equal(O.safe_sign_fwd(v), z["bcd_fwd"]) and np.array_equal(O.safe_sign_bwd(g, v), z["bcd_bwd"])
    assert np.array_equal(O.safe_sign_fwd(v), z["safe_sign"])


def test_gf_blend_oracle_matches_reference_fixture():
    """with_gf=True (resnet_EE.py:185-191): Gaussian on the edge map, blend, and the autograd through both."""
    z = np.load(os.path.join(GOLD, "gf_blend.npz"))
    w = float(z["w"])
    out = O.gf_blend_fwd(z["edge"], z["base"], w)
    np.testing.assert_allclose(out, z["out"], rtol=1e-5, atol=1e-7)
    g_edge, g_base = O.gf_blend_bwd(z["g_out"], z["edge"], z["base"], w)
    assert np.array_equal(g_base, z["g_base"])
    assert np.abs(g_edge - z["g_edge"]).max() <= 1e-5 * np.abs(z["g_edge"]).max()
    # and the whole chain: filter backward of that g_edge == the reference's input gradient
    p = O.make_params("step125", high=float(z["high"]))
    assert np.array_equal(O.edge_fwd(z["x"], p), z["edge"])
    g_x = O.edge_bwd(g_edge, z["x"], p)
    fin = np.isfinite(z["g_x"])
    assert np.abs(g_x - z["g_x"])[fin].max() <= 1e-5 * np.abs(z["g_x"][fin]).max()


def head_gradient(zt, y, weight):
    """dL/dz of the fixture model's head (oracle/make_golden.py TinyEENet): logits = z.flatten(1) @ W^T, CE(sum)."""
    import torch
    z = torch.from_numpy(zt).requires_grad_()
    logits = z.reshape(z.shape[0], -1) @ torch.from_numpy(weight).t()
    loss = torch.nn.functional.cross_entropy(logits, torch.from_numpy(y), reduction='sum')
    return torch.autograd.grad(loss, [z])[0].numpy()


@pytest.mark.parametrize("variant", ["step125", "canny", "bpda"])
def test_oracle_teacher_forced_pgd10(variant):
    """The reference's own PGD-10 run (utils/attacks.py:12-29), replayed teacher-forced: at every reference iterate x_i
    the edge mask is exact and the input gradient within 1e-5 of max |g|; the step applied to the REFERENCE's (x_i, g_i)
    reproduces the reference's next iterate, so the last one equals its x_adv bit for bit."""
    import torch
    torch.set_num_threads(1)
    z = np.load(os.path.join(GOLD, "pgd10_traced_%s.npz" % variant))
    x0, y, gs = z["x"], z["y"], z["gs"]
    B, C, H, W = x0.shape
    steps = gs.shape[0]
    edges = np.unpackbits(z["edges"])[:steps * B * H * W].reshape(steps, B, 1, H, W).astype(np.float32)
    weight = (np.random.default_rng(int(z["head_seed"])).standard_normal((int(z["n_class"]), C * H * W)).astype(np.float32) * 0.05)
    p = O.make_params(variant, low=38 / 255, high=76 / 255, hysteresis=True)
    x = x0.copy()
    for i in range(steps):
        out, edge = O.edge_blend_fwd(x, x, p, 1.0, want_edge=True)
        assert np.array_equal(edge, edges[i]), "iteration %d: %d mask pixels differ" % (i, (edge != edges[i]).sum())
        g_z = head_gradient(out, y, weight)
        g_x, g_base = O.edge_blend_bwd(g_z, x, x, p, 1.0)
        g = g_x + g_base
        assert np.abs(g - gs[i]).max() <= 1e-5 * np.abs(gs[i]).max(), (i, np.abs(g - gs[i]).max(), np.abs(gs[i]).max())
        x = O.pgd_linf_step(x, gs[i], x0, 2 / 255, 16 / 255)              # teacher-forced: the reference's gradient
    assert np.array_equal(x, z["x_adv"])
