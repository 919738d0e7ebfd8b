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
