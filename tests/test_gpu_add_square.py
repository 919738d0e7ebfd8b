"""-m gpu: Add_Square (utils/core.py:589-655) -- the fused CUDA kernels against the
oracle, the reference fixtures, and the seeded drop-in module (same random draws as the reference)."""
import glob
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from edge_enhancement_b200 import core, functional as F_ee   # noqa: E402
from oracle import oracle as O                               # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SQUARE_FILES = sorted(glob.glob(os.path.join(GOLD, "add_square_*.npz")))
DEV = "cuda:0"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.mark.parametrize("path", SQUARE_FILES, ids=lambda p: os.path.basename(p)[11:-4])
def test_kernels_match_reference_fixture(path):
    z = np.load(path)
    eps, nq = float(str(z["meta"][2])), int(str(z["meta"][3]))
    out = F_ee.add_square(cu(z["x"]), cu(z["stripe"]), cu(z["table"]), eps)
    assert np.array_equal(out.cpu().numpy(), z["out"])                         # forward: bit-exact vs the reference
    g_x = F_ee.add_square_backward(cu(z["g"]), cu(z["x"]), cu(z["stripe"]), cu(z["table"]), eps).cpu().numpy()
    assert np.array_equal(g_x, O.add_square(z["x"], z["stripe"], z["table"], eps, g=z["g"]))   # == oracle, bits
    if nq == 1:
        assert np.array_equal(g_x, z["g_x"])
    else:
        np.testing.assert_allclose(g_x, z["g_x"], rtol=2e-7, atol=0)


@pytest.mark.parametrize("path", SQUARE_FILES, ids=lambda p: os.path.basename(p)[11:-4])
def test_seeded_module_reproduces_the_reference_run(path):
    """Drop-in property: with the generator seeded like the reference run, the module draws the same stripes and
    squares and returns the same tensor; gradients flow through the autograd.Function."""
    z = np.load(path)
    C, S, eps, nq, seed = int(str(z["meta"][0])), int(str(z["meta"][1])), float(str(z["meta"][2])), int(str(z["meta"][3])), int(str(z["meta"][4]))
    m = core.Add_Square(channels=C, size=S, epsilon=eps, n_queries=nq)
    x = cu(z["x"]).requires_grad_()
    torch.manual_seed(seed)
    out = m(x)
    assert np.array_equal(out.detach().cpu().numpy(), z["out"])
    out.backward(cu(z["g"]))
    np.testing.assert_allclose(x.grad.cpu().numpy(), z["g_x"], rtol=2e-7, atol=0)


@pytest.mark.parametrize("shape,nq", [((5, 3, 32, 32), 1), ((2, 1, 28, 28), 3), ((3, 2, 17, 17), 2), ((64, 3, 64, 64), 1)])
def test_kernels_match_oracle_random(shape, nq):
    B, C, H, W = shape
    r = np.random.default_rng(B * 1000 + H)
    eps = np.float32(0.07)
    x = r.random(shape, dtype=np.float32)
    x.reshape(-1)[::5] = 0.0
    x.reshape(-1)[2::9] = 1.0
    g = r.standard_normal(shape, dtype=np.float32)
    stripe = np.sign(r.random((B, C, W), dtype=np.float32) * 2 - 1).astype(np.float32)
    stripe.reshape(-1)[::13] = 0.0                                             # sign(0) = 0 is possible in the reference
    table = np.zeros((nq, 2 + C), np.float32)
    for q in range(nq):
        s = int(r.integers(1, H + 1))
        table[q, 0], table[q, 1] = int(r.integers(0, H - s + 1)), s
        table[q, 2:] = 2 * eps * np.sign(r.random(C) - 0.5)
    out = F_ee.add_square(cu(x), cu(stripe), cu(table), float(eps)).cpu().numpy()
    assert np.array_equal(out, O.add_square(x, stripe, table, eps))
    g_x = F_ee.add_square_backward(cu(g), cu(x), cu(stripe), cu(table), float(eps)).cpu().numpy()
    assert np.array_equal(g_x, O.add_square(x, stripe, table, eps, g=g))
    assert (np.abs(out - x) <= eps + 1e-7).all() and out.min() >= 0 and out.max() <= 1


def test_module_errors():
    m = core.Add_Square(channels=3, size=32, epsilon=0.05, n_queries=1)
    with pytest.raises(RuntimeError):
        m(torch.rand(2, 3, 32, 32))                       # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        m(torch.rand(2, 3, 16, 16, device=DEV))           # wrong size
