"""not gpu, build container only: the oracle against the LIVE reference modules on fresh random
inputs (skipped where /root/reference is absent, e.g. on the GPU box)."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from oracle import ref_loader
from tests import common as T

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ref():
    return ref_loader.load()


CASES = [("step125", "CannyFilter_step125_1"), ("canny", "CannyFilter"), ("bpda", "CannyFilter_BPDA")]


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("shape", [(2, 3, 32, 32), (3, 1, 28, 28), (1, 3, 13, 10)], ids=str)
@pytest.mark.parametrize("variant,cls", CASES)
def test_against_live_reference(ref, variant, cls, shape, seed):
    rc, _ = ref
    x, base, g_out, _ = T.make_inputs(1000 * seed + shape[2], *shape)
    with ref_loader.quiet():
        f = getattr(rc, cls)(use_cuda=False, alpha=0.02)
    xt = torch.from_numpy(x).requires_grad_()
    bt = torch.from_numpy(base).requires_grad_()
    e = f(xt, low_threshold=T.LOW, high_threshold=T.HIGH, hysteresis=True)
    out = torch.clamp(bt + 1.0 * e, 0, 1)
    out.backward(torch.from_numpy(g_out))
    p = O.make_params(variant, alpha=0.02, low=T.LOW, high=T.HIGH, hysteresis=True)
    o_out, o_edge = O.edge_blend_fwd(x, base, p, 1.0, want_edge=True)
    o_gx, o_gb = O.edge_blend_bwd(g_out, x, base, p, 1.0)
    mism = (o_edge != e.detach().numpy())
    if mism.any():          # a flip is only legitimate within a few ulp of a threshold / NMS tie
        pytest.fail("%d mask pixels differ from the live reference" % mism.sum())
    np.testing.assert_allclose(o_out, out.detach().numpy(), rtol=1e-5, atol=1e-7)
    ref_g = xt.grad.numpy()
    fin = np.isfinite(ref_g)
    assert np.abs(o_gx - ref_g)[fin].max() <= 1e-5 * np.abs(ref_g[fin]).max()
    assert np.array_equal(o_gb, bt.grad.numpy())


def test_attack_expressions_live(ref):
    _, ra = ref
    x, g, x0 = T.make_attack_inputs(77, (3, 3, 9, 11), 16 / 255)
    tx, tg, tx0 = map(torch.from_numpy, (x, g, x0))
    for a in (2 / 255, -2 / 255):
        want = torch.clamp(torch.min(torch.max(tx + a * torch.sign(tg), tx0 - 16 / 255), tx0 + 16 / 255), 0, 1)
        assert np.array_equal(O.pgd_linf_step(x, g, x0, a, 16 / 255), want.numpy())
    n = ra.l2_norm(tg).numpy()
    assert np.allclose(n, np.sqrt((g.reshape(3, -1) ** 2).mean(1)), rtol=1e-6)


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE sizes (M = 128x1x28x28 sparse, T = 256x3x64x64, I = 32x3x224x224), tie-aware.
#
# The binary mask is a threshold / strict comparison on fp32 values whose summation order oneDNN does not fix: the live
# reference disagrees WITH ITSELF between batch 128 and batch 1 on exact NMS ties (an isolated impulse on a flat MNIST
# background has mathematically equal neighbours; which one is 1 ulp larger depends on the convolution's blocking).  So a
# mismatch between the oracle and the batched reference is accepted only if
#   (a) the reference run ONE IMAGE AT A TIME gives the oracle's value at that pixel, or
#   (b) a float64 recomputation shows that a deciding comparison (NMS neighbour, low / high threshold, alpha gate, or the
#       half-to-even rounding boundary of the orientation bin, core.py:258-260) has a margin of at most 4 fp32 ulp
#       (8 ulp for the orientation, which sits behind an atan and three more roundings).
# Everything else fails.  The counts and margins are printed (pytest -s) and recorded in DESIGN.md section 3.
# ---------------------------------------------------------------------------------------------------------------------
def _f64_magnitude(x, sigma=1.0):
    """gated-free gradient magnitude of one image [C,H,W] in float64 (core.py:560-571 / :243-256)."""
    g = O.gaussian3(0.0, sigma).astype(np.float64)
    kx = np.array([[-0.5, 0, 0.5], [-1, 0, 1], [-0.5, 0, 0.5]])

    def corr(p, k):
        H, W = p.shape[0] - 2, p.shape[1] - 2
        return sum(k[a, b] * p[a:a + H, b:b + W] for a in range(3) for b in range(3))
    gx = gy = 0.0
    for c in range(x.shape[0]):
        bl = corr(np.pad(x[c].astype(np.float64), 1, mode="edge"), g)
        pb = np.pad(bl, 1, mode="edge")
        gx = gx + corr(pb, kx)
        gy = gy + corr(pb, kx.T)
    gx, gy = gx / x.shape[0], gy / x.shape[0]
    with np.errstate(divide="ignore", invalid="ignore"):
        ori45 = (np.arctan(gy / gx) * (360.0 / np.pi) + 180.0) / 45.0          # core.py:258-260: rounded half-to-even
    return np.sqrt(gx * gx + gy * gy), ori45


def _decision_margin_ulp(mag64, ori45, r, c, thresholds, with_orientation):
    """smallest distance, in fp32 ulp, between a value the filter compares at (r, c) and what it is compared with: the
    magnitude against its 8 neighbours (NMS) and the thresholds, and the orientation against its rounding boundary."""
    H, W = mag64.shape
    m = mag64[r, c]
    ulp = float(np.spacing(np.float32(max(m, 1e-30))))
    cands = [abs(m - t) / ulp for t in thresholds if t is not None]
    for dr in (-1, 0, 1):
        for dc in (-1, 0, 1):
            if (dr or dc) and 0 <= r + dr < H and 0 <= c + dc < W:
                cands.append(abs(m - mag64[r + dr, c + dc]) / ulp)
    v = ori45[r, c]
    if with_orientation and np.isfinite(v):
        cands.append(abs(v - (np.floor(v) + 0.5)) / float(np.spacing(np.float32(v))))
    return min(cands)


def _tie_map(mag64, ori45, thresholds, with_orientation):
    """boolean map of the pixels whose NMS / threshold / orientation decision is within the tie margins (vectorised
    version of _decision_margin_ulp over one image)."""
    ulp = np.spacing(np.maximum(mag64, 1e-30).astype(np.float32)).astype(np.float64)
    pad = np.pad(mag64, 1, mode="constant", constant_values=np.inf)
    H, W = mag64.shape
    best = np.full_like(mag64, np.inf)
    for dr in (0, 1, 2):
        for dc in (0, 1, 2):
            if dr != 1 or dc != 1:
                best = np.minimum(best, np.abs(mag64 - pad[dr:dr + H, dc:dc + W]))
    for t in thresholds:
        if t is not None:
            best = np.minimum(best, np.abs(mag64 - t))
    tie = best / ulp <= 4.0
    if with_orientation:
        v = np.where(np.isfinite(ori45), ori45, 0.25)
        tie |= np.abs(v - (np.floor(v) + 0.5)) / np.spacing(v.astype(np.float32)).astype(np.float64) <= 8.0
    return tie


def _dilate(m, k):
    out = m.copy()
    for _ in range(k):
        p = np.pad(out, 1)
        out = np.zeros_like(m)
        for dr in (0, 1, 2):
            for dc in (0, 1, 2):
                out |= p[dr:dr + m.shape[0], dc:dc + m.shape[1]]
    return out


BASELINE_CASES = [
    # id, variant, class, shape, input kind, alpha, low, high
    ("M-canny", "canny", "CannyFilter", (128, 1, 28, 28), "sparse", 0.3, 25 / 255, 51 / 255),
    ("M-bpda", "bpda", "CannyFilter_BPDA", (128, 1, 28, 28), "sparse", 0.0, 25 / 255, 51 / 255),
    ("M-step125", "step125", "CannyFilter_step125_1", (128, 1, 28, 28), "sparse", 0.3, None, 51 / 255),
    ("T-canny", "canny", "CannyFilter", (256, 3, 64, 64), "uniform", 0.0, 38 / 255, 76 / 255),
    ("T-bpda", "bpda", "CannyFilter_BPDA", (256, 3, 64, 64), "uniform", 0.0, 38 / 255, 76 / 255),
    ("T-step125", "step125", "CannyFilter_step125_1", (256, 3, 64, 64), "uniform", 0.0, None, 76 / 255),
    ("I-canny", "canny", "CannyFilter", (32, 3, 224, 224), "uniform", 0.0, 38 / 255, 76 / 255),
    ("I-step125", "step125", "CannyFilter_step125_1", (32, 3, 224, 224), "uniform", 0.0, None, 76 / 255),
]


@pytest.mark.parametrize("seed", [3, 11])
@pytest.mark.parametrize("case", BASELINE_CASES, ids=lambda c: c[0])
def test_baseline_sizes_tie_aware(ref, case, seed):
    rc, _ = ref
    cid, variant, cls, shape, kind, alpha, low, high = case
    if shape[2] == 224 and seed != 3:
        pytest.skip("one seed at the ImageNet size keeps the CPU suite short")
    x, base, g_out, _ = T.make_inputs(seed, *shape, kind=kind)
    with ref_loader.quiet():
        f = getattr(rc, cls)(use_cuda=False, alpha=alpha)
    xt = torch.from_numpy(x).requires_grad_()
    bt = torch.from_numpy(base).requires_grad_()
    e = f(xt, low_threshold=low, high_threshold=high, hysteresis=True)
    out = torch.clamp(bt + 1.0 * e, 0, 1)
    out.backward(torch.from_numpy(g_out))
    ref_edge = e.detach().numpy()
    p = O.make_params(variant, alpha=alpha, low=low, high=high, hysteresis=True)
    o_out, o_edge = O.edge_blend_fwd(x, base, p, 1.0, want_edge=True)
    o_gx, o_gb = O.edge_blend_bwd(g_out, x, base, p, 1.0)

    mism = np.argwhere(o_edge != ref_edge)
    by_single, by_margin, margins, unexplained = 0, 0, [], []
    single_cache = {}
    for (b, _, r, c) in mism:
        if b not in single_cache:
            with torch.no_grad():
                single_cache[b] = f(torch.from_numpy(x[b:b + 1]), low_threshold=low, high_threshold=high, hysteresis=True).numpy()[0, 0]
        if single_cache[b][r, c] == o_edge[b, 0, r, c]:
            by_single += 1
            continue
        # hysteresis looks at the 3x3 neighbourhood: the tie may sit on a neighbour
        mag64, ori45 = _f64_magnitude(x[b])
        thr = (low, high, alpha if (variant != "bpda" and alpha > 0) else None)
        nbrs = [(rr, cc) for rr in range(max(r - 1, 0), min(r + 2, shape[2])) for cc in range(max(c - 1, 0), min(c + 2, shape[3]))]
        mg = min(_decision_margin_ulp(mag64, ori45, rr, cc, thr, False) for rr, cc in nbrs)
        mo = min(_decision_margin_ulp(mag64, ori45, rr, cc, (), variant != "step125") for rr, cc in nbrs)
        margins.append(min(mg, mo))
        if mg <= 4.0 or mo <= 8.0:
            by_margin += 1
        else:
            unexplained.append((int(b), int(r), int(c), float(mg), float(mo)))
    print("\n[tie-aware] %-10s seed %2d: %d / %d mask pixels differ from the batched reference; %d equal the reference run one "
          "image at a time, %d within the tie margins (4 ulp NMS / threshold, 8 ulp orientation; smallest margins %s), %d unexplained"
          % (cid, seed, len(mism), o_edge.size, by_single, by_margin, ["%.2f" % m for m in margins[:6]], len(unexplained)))
    assert not unexplained, "mask mismatches that are neither reference non-determinism nor <= 4 ulp ties: %s" % unexplained[:5]
    # a handful of exact ties per batch is the expected worst case (MNIST-like inputs); the rest must be identical
    assert len(mism) <= max(8, o_edge.size // 20000)
    # everything downstream of the mask, on the images whose masks agree
    ok = np.ones(shape[0], bool)
    ok[np.unique(mism[:, 0])] = False if len(mism) else True
    np.testing.assert_allclose(o_out[ok], out.detach().numpy()[ok], rtol=1e-5, atol=1e-7)
    assert np.array_equal(o_gb[ok], bt.grad.numpy()[ok])
    ref_g = xt.grad.numpy()
    fin = np.isfinite(ref_g) & ok[:, None, None, None]
    assert fin.any()
    tol = 1e-5 * np.abs(ref_g[fin]).max()
    dev = fin & (np.abs(o_gx - ref_g) > tol)
    # The gradient also depends on NMS decisions of sub-threshold pixels that never show in the mask (CannyFilter passes
    # gradient through every non-removed pixel): a deviation is accepted only within the 5 x 5 footprint of the two adjoint
    # stencils around a pixel whose NMS / orientation decision is a tie by the same margins as above.
    n_dev = 0
    for b in np.unique(np.argwhere(dev)[:, 0]) if dev.any() else []:
        mag64, ori45 = _f64_magnitude(x[b])
        thr = (low, high, alpha if (variant != "bpda" and alpha > 0) else None)
        near_tie = _dilate(_tie_map(mag64, ori45, thr, variant != "step125"), 2)
        stray = dev[b].any(axis=0) & ~near_tie
        assert not stray.any(), "%s image %d: %d gradient deviations away from any tie" % (cid, b, stray.sum())
        n_dev += int(dev[b].any(axis=0).sum())
    print("[tie-aware] %-10s seed %2d: g_x within 1e-5 of max|g| except %d pixel(s) next to an NMS / orientation tie" % (cid, seed, n_dev))
    assert n_dev <= max(4, o_edge.size // 50000)
