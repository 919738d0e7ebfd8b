"""-m gpu: native HighFreqSuppress kernel (ee_hfs_f32) against the C oracle (bit for bit: same
tables, same fmaf chains), against the torch.fft restatement (1e-5; the reference's own torch.rfft version cannot run,
so parity with the reference is UNPINNED), and through the drop-in module incl. autograd (the operator is symmetric)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from edge_enhancement_b200 import core, functional as F_ee   # noqa: E402
from oracle import oracle as O                               # noqa: E402

DEV = "cuda:0"
SUPPORTED = [(64, 8), (28, 4), (32, 8), (224, 16), (128, 12), (288, 18)]


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.mark.parametrize("N,r", SUPPORTED)
@pytest.mark.parametrize("planes", [(1, 1), (2, 3), (37, 3), (5, 1)], ids=str)
def test_kernel_matches_oracle_bit_for_bit(N, r, planes):
    B, C = planes
    x = np.random.default_rng(N + B).standard_normal((B, C, N, N)).astype(np.float32)
    x[0, 0, :3] = 0.0
    assert F_ee.hfs_supported(N, r)
    got = F_ee.hfs(cu(x), r).cpu().numpy()
    assert np.array_equal(got, O.hfs(x, r))


@pytest.mark.parametrize("N,r", SUPPORTED)
def test_module_matches_torch_fft_and_autograd(N, r):
    m = core.HighFreqSuppress(N, N, r)
    gen = torch.Generator(device=DEV).manual_seed(N)
    x = torch.rand((6, 3, N, N), device=DEV, generator=gen, requires_grad=True)
    g = torch.randn((6, 3, N, N), device=DEV, generator=gen)
    y = m(x)                                   # native kernel (autograd.Function)
    y.backward(g)
    x2 = x.detach().clone().requires_grad_()
    y2 = m._fft_forward(x2)                    # torch.fft restatement
    y2.backward(g)
    scale = float(y2.abs().max())
    assert float((y - y2).abs().max()) <= 1e-5 * scale
    assert float((x.grad - x2.grad).abs().max()) <= 1e-5 * float(x2.grad.abs().max())
    # low-pass sanity: the mean passes, a checkerboard (Nyquist) is removed
    assert torch.allclose(y.mean((2, 3)), x.mean((2, 3)), atol=1e-5)
    cb = ((torch.arange(N, device=DEV)[:, None] + torch.arange(N, device=DEV)[None, :]) % 2).float()[None, None]
    assert float(m(cb - 0.5).abs().max()) < 1e-5


def test_unsupported_shapes_raise_and_non_contiguous_inputs_work():
    assert not F_ee.hfs_supported(96, 12)
    m = core.HighFreqSuppress(96, 96, 12)
    x = torch.rand((2, 3, 96, 96), device=DEV)
    with pytest.raises(RuntimeError):
        m(x)                                     # no silent torch.fft fallback
    assert torch.equal(core.HighFreqSuppress(96, 96, 12, impl='torch_fft')(x), m._fft_forward(x))
    with pytest.raises(RuntimeError):
        F_ee.hfs(x, 12)
    m64 = core.HighFreqSuppress(64, 64, 8)
    xt = torch.rand((2, 64, 64, 3), device=DEV).permute(0, 3, 1, 2)           # channels_last view: copied to NCHW planes
    assert float((m64(xt) - m64._fft_forward(xt)).abs().max()) < 1e-5


@pytest.mark.parametrize("N,r", [(64, 8), (28, 4), (224, 16)])
def test_accumulate_mode_matches_oracle(N, r):
    rng = np.random.default_rng(3)
    x = rng.standard_normal((4, 3, N, N)).astype(np.float32)
    add = rng.standard_normal((4, 3, N, N)).astype(np.float32)
    want = O.hfs(x, r, add=add)
    assert np.array_equal(F_ee.hfs(cu(x), r, add=cu(add)).cpu().numpy(), want)
    buf = cu(add)                                         # in place: out aliases add
    F_ee.hfs(cu(x), r, out=buf, add=buf)
    assert np.array_equal(buf.cpu().numpy(), want)


@pytest.mark.parametrize("cize,r,variant", [(64, 8, "CannyFilter_step125_1"), (64, 8, "CannyFilter"), (28, 4, "CannyFilter_step125_1")])
def test_fused_front_end_equals_the_three_separate_nodes(cize, r, variant):
    """core.EdgeEnhance on a supported shape runs low-pass + edge filter + blend as ONE autograd node (the low-pass of the
    backward accumulates into the edge-path gradient inside the kernel); results and gradients are identical to composing
    HighFreqSuppress, the filter and the blend as separate nodes."""
    import contextlib, io
    C = 1 if cize == 28 else 3
    with contextlib.redirect_stdout(io.StringIO()):
        m = core.EdgeEnhance(cize=cize, r=r, w=1.0, low=38.0, high=76.0, alpha=0.0, sigma=1, type_canny=variant)
    gen = torch.Generator(device=DEV).manual_seed(cize)
    x = torch.rand((8, C, cize, cize), device=DEV, generator=gen, requires_grad=True)
    g = torch.randn((8, C, cize, cize), device=DEV, generator=gen)
    y = m(x)
    y.backward(g)
    x2 = x.detach().clone().requires_grad_()
    base = m.hfs(x2)
    y2 = core.edge_enhance(x2, base, m.canny, m.w, m.low, m.high, True)
    y2.backward(g)
    assert torch.equal(y, y2)
    assert torch.equal(x.grad, x2.grad)


@pytest.mark.parametrize("B,N,r", [(2048, 64, 8), (8192, 28, 4), (176, 224, 16)])
def test_persistent_loop_many_planes(B, N, r):
    """More plane groups than resident CTAs: every CTA walks several groups while the next planes stream in behind it
    (cp.async into the buffer stage 1 has just released).  Bit-exact against the oracle on planes from the first, a
    middle and the last round, and 1e-5 against the torch.fft restatement on everything; repeated launches agree."""
    C = 1 if N == 28 else 3
    gen = torch.Generator(device=DEV).manual_seed(B)
    x = torch.randn((B, C, N, N), device=DEV, generator=gen)
    y = F_ee.hfs(x, r)
    ref = core.HighFreqSuppress(N, N, r)._fft_forward(x)
    assert float((y - ref).abs().max()) <= 1e-5 * float(ref.abs().max())
    for lo in (0, B // 2 - 1, B - 3):
        xs = x[lo:lo + 3].cpu().numpy()
        assert np.array_equal(y[lo:lo + 3].cpu().numpy(), O.hfs(xs, r)), lo
    assert torch.equal(F_ee.hfs(x, r), y)


# ---- the tensor-core variant (ee_hfs_tc_f32: tcgen05.mma kind::tf32, 3 x TF32 split) -------------------------------------------
# Floating point on an unspecified accumulation order: NOT bit-identical to the oracle.  Tolerance 4e-6 * max|y| against the
# oracle (measured 1.5e-6 .. 1.7e-6 on [0,1] inputs at 4096x3x64x64; the oracle itself is 0.6e-6 from float64).
TC_TOL = 4e-6


@pytest.mark.parametrize("planes", [1, 2, 3, 5, 64, 593, 12288])
@pytest.mark.parametrize("kind", ["uniform", "normal"])
def test_tensor_core_kernel_matches_oracle_within_tolerance(planes, kind):
    rng = np.random.default_rng(planes)
    x = (rng.random((planes, 64, 64)) if kind == "uniform" else rng.standard_normal((planes, 64, 64))).astype(np.float32)
    x[0, :3] = 0.0
    assert F_ee.hfs_supported(64, 8, 'tcgen05')
    xd = cu(x)
    got = F_ee.hfs(xd, 8, impl='tcgen05')
    ref = F_ee.hfs(xd, 8)                               # == oracle bit for bit (test above); the oracle itself for small counts
    if planes <= 64:
        assert np.array_equal(ref.cpu().numpy(), O.hfs(x, 8))
    err = float((got - ref).abs().max())
    assert err <= TC_TOL * float(ref.abs().max()), err
    assert not bool(torch.isnan(got).any())


def test_tensor_core_kernel_accumulate_mode_and_float64_error():
    gen = torch.Generator(device=DEV).manual_seed(5)
    x = torch.rand((1000, 64, 64), device=DEV, generator=gen)
    add = torch.randn((1000, 64, 64), device=DEV, generator=gen)
    m = core.HighFreqSuppress(64, 64, 8, impl='torch_fft')
    exact = m._fft_forward(x.double())
    got = F_ee.hfs(x, 8, impl='tcgen05')
    e_tc = float((got.double() - exact).abs().max())
    e_ffma = float((F_ee.hfs(x, 8).double() - exact).abs().max())
    print("max abs error against float64: tcgen05 %.3e, FFMA %.3e" % (e_tc, e_ffma))
    assert e_tc <= 3e-6 and e_ffma <= 1.5e-6
    out = add.clone()
    F_ee.hfs(x, 8, out=out, add=out, impl='tcgen05')    # y = H x + add, in place
    assert float((out - (got + add)).abs().max()) <= 1e-6
    assert F_ee.hfs(x[:0], 8, impl='tcgen05').shape == (0, 64, 64)


def test_tensor_core_module_autograd_and_unsupported_shapes():
    m = core.HighFreqSuppress(64, 64, 8, impl='tcgen05')
    gen = torch.Generator(device=DEV).manual_seed(9)
    x = torch.rand((7, 3, 64, 64), device=DEV, generator=gen, requires_grad=True)
    g = torch.randn((7, 3, 64, 64), device=DEV, generator=gen)
    y = m(x)
    y.backward(g)
    x2 = x.detach().clone().requires_grad_()
    y2 = m._fft_forward(x2)
    y2.backward(g)
    assert float((y - y2).abs().max()) <= 1e-5 * float(y2.abs().max())
    assert float((x.grad - x2.grad).abs().max()) <= 1e-5 * float(x2.grad.abs().max())
    assert not F_ee.hfs_supported(28, 4, 'tcgen05')
    with pytest.raises(RuntimeError):
        core.HighFreqSuppress(28, 28, 4, impl='tcgen05')(torch.rand((2, 1, 28, 28), device=DEV))
    with pytest.raises(RuntimeError):
        F_ee.hfs(torch.rand((2, 28, 28), device=DEV), 4, impl='tcgen05')


def test_front_end_with_the_tensor_core_low_pass_and_in_place_accumulation():
    """EdgeEnhance(hfs_impl='tcgen05'): same masks, blended image and gradient within the tolerances of north_star (1e-5); the
    backward accumulates H(g_base) into g_x through the TMA reduction store (add aliasing y)."""
    import contextlib, io
    gen = torch.Generator(device=DEV).manual_seed(21)
    x = torch.rand((37, 3, 64, 64), device=DEV, generator=gen)
    g = torch.randn((37, 3, 64, 64), device=DEV, generator=gen)
    outs = []
    for impl in ('native', 'tcgen05'):
        with contextlib.redirect_stdout(io.StringIO()):
            m = core.EdgeEnhance(cize=64, r=8, w=1.0, low=38.0, high=76.0, type_canny='CannyFilter_step125_1', hfs_impl=impl).to(DEV)
        xi = x.clone().requires_grad_()
        y = m(xi)
        y.backward(g)
        outs.append((y.detach(), xi.grad.detach()))
    (y0, g0), (y1, g1) = outs
    assert float((y0 - y1).abs().max()) <= 1e-5
    assert float((g0 - g1).abs().max()) <= 1e-5 * float(g0.abs().max())
    # the in-place accumulation alone, odd plane count: y = H x + y
    acc = torch.randn((5, 64, 64), device=DEV, generator=gen)
    xs = torch.rand((5, 64, 64), device=DEV, generator=gen)
    want = F_ee.hfs(xs, 8) + acc
    got = acc.clone()
    F_ee.hfs(xs, 8, out=got, add=got, impl='tcgen05')
    assert float((got - want).abs().max()) <= TC_TOL * float(want.abs().max())


def test_tensor_core_kernel_is_cuda_graph_capturable():
    """The tensor maps are encoded on the host and passed by value: the call can be captured and replayed on new contents."""
    gen = torch.Generator(device=DEV).manual_seed(3)
    x = torch.rand((48, 64, 64), device=DEV, generator=gen)
    y = torch.empty_like(x)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        F_ee.hfs(x, 8, out=y, impl='tcgen05')             # warm-up outside the capture (attribute set-up)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        F_ee.hfs(x, 8, out=y, impl='tcgen05')
    x.copy_(torch.rand((48, 64, 64), device=DEV, generator=gen))
    graph.replay()
    torch.cuda.synchronize()
    ref = F_ee.hfs(x, 8)
    assert float((y - ref).abs().max()) <= TC_TOL * float(ref.abs().max())
