"""-m gpu: BASELINE.json's full sizes.  Where the oracle finishes in seconds the comparison is direct
(bit-exact); at sweep sizes it goes through size-independent properties: strip / kernel-choice
invariance, batch-permutation equivariance, channel-identical gradients, mask consistency of g_base,
range / idempotence of the projection."""
import numpy as np
import pytest
import torch

from tests import common as T

pytestmark = pytest.mark.gpu

from edge_enhancement_b200 import functional as F_ee, _lib   # noqa: E402
from oracle import oracle as O                               # noqa: E402

DEV = "cuda:0"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.fixture(autouse=True)
def _reset():
    _lib.load().ee_set_tuning(0, 0, 0)
    yield
    _lib.load().ee_set_tuning(0, 0, 0)


FULL = [
    # config, variant, shape, alpha, low, high
    ("M: MNIST 128x1x28x28 full Canny (ee_at_training.yml)", "canny", (128, 1, 28, 28), 0.3, 25 / 255, 51 / 255),
    ("M: MNIST 128x1x28x28 step125", "step125", (128, 1, 28, 28), 0.0, None, 51 / 255),
    ("T: Tiny 256x3x64x64 step125 (ee_at_bpda3_square.yml)", "step125", (256, 3, 64, 64), 0.0, None, 76 / 255),
    ("T: Tiny 256x3x64x64 full Canny (ee_at_training.yml)", "canny", (256, 3, 64, 64), 0.0, 38 / 255, 76 / 255),
    ("T: Tiny 256x3x64x64 BPDA", "bpda", (256, 3, 64, 64), 0.0, 38 / 255, 76 / 255),
    ("I: ImageNet per-GPU 32x3x224x224 step125 (at_ee_training.yml)", "step125", (32, 3, 224, 224), 0.0, None, 76 / 255),
]


@pytest.mark.parametrize("cfg", FULL, ids=lambda c: c[0].split(":")[0] + "-" + c[1])
def test_full_size_configs_bit_exact(cfg):
    _, variant, shape, alpha, low, high = cfg
    kind = "sparse" if shape[1] == 1 else "uniform"
    x, base, g_out, _ = T.make_inputs(42, *shape, kind=kind)
    g = O.gaussian3()
    pc = F_ee.make_params(variant, g, alpha, low, high, True)
    po = O.make_params(variant, alpha=alpha, low=low, high=high, hysteresis=True)
    out = F_ee.edge_blend(cu(x), cu(base), pc, 1.0)
    g_x, g_base = F_ee.edge_blend_backward(cu(g_out), cu(x), cu(base), pc, 1.0)
    o_out = O.edge_blend_fwd(x, base, po, 1.0)
    o_gx, o_gb = O.edge_blend_bwd(g_out, x, base, po, 1.0)
    assert np.array_equal(out.cpu().numpy(), o_out)
    assert np.array_equal(g_base.cpu().numpy(), o_gb)
    assert np.array_equal(g_x.cpu().numpy(), o_gx)


@pytest.mark.parametrize("variant", ["step125", "canny"])
@pytest.mark.parametrize("shape", [(1024, 3, 64, 64), (192, 3, 224, 224), (2048, 3, 32, 32)], ids=str)
def test_sweep_size_properties(variant, shape):
    B, C, H, W = shape
    gen = torch.Generator(device=DEV).manual_seed(7)
    x = torch.rand(shape, device=DEV, generator=gen)
    base = torch.rand(shape, device=DEV, generator=gen) * 1.1 - 0.1
    g_out = torch.randn(shape, device=DEV, generator=gen)
    low = None if variant == "step125" else T.LOW
    p = F_ee.make_params(variant, O.gaussian3(), 0.0, low, T.HIGH, True)
    L = _lib.load()

    def run():
        out, edge = F_ee.edge_blend(x, base, p, 1.0, want_edge=True)
        g_x, g_base = F_ee.edge_blend_backward(g_out, x, base, p, 1.0)
        return out, edge, g_x, g_base

    ref = run()
    # (1) the result does not depend on the strip height nor on which kernel family runs
    for th, staging in ((5, 0), (16, 1), (11, 4)):
        L.ee_set_tuning(th, th, staging)
        for a, b in zip(ref, run()):
            assert torch.equal(a, b), (th, staging)
    L.ee_set_tuning(0, 0, 0)
    out, edge, g_x, g_base = ref
    # (2) images are independent: permuting the batch permutes the result
    perm = torch.randperm(B, device=DEV, generator=gen)
    out_p = F_ee.edge_blend(x[perm].contiguous(), base[perm].contiguous(), p, 1.0)
    assert torch.equal(out_p, out[perm])
    # (3) masks are binary, the blend is in range, every channel receives the same edge gradient
    assert bool(((edge == 0) | (edge == 1)).all()) and 0.02 < float(edge.mean()) < 0.6
    assert float(out.min()) >= 0 and float(out.max()) <= 1
    assert torch.equal(out, torch.clamp(base + 1.0 * edge, 0, 1))              # blend identity, bit-exact
    for c in range(1, C):
        assert torch.equal(g_x[:, 0], g_x[:, c])
    assert bool(torch.isfinite(g_x).all())
    # (4) g_base is g_out under the inclusive clamp mask
    pre = base + 1.0 * edge
    mask = (pre >= 0) & (pre <= 1)
    assert torch.equal(g_base, torch.where(mask, g_out, torch.zeros_like(g_out)))
    # (5) linearity of the adjoint in the upstream gradient (exact for a power-of-two scale)
    g_x2, g_base2 = F_ee.edge_blend_backward(2.0 * g_out, x, base, p, 1.0)
    assert torch.equal(g_x2, 2.0 * g_x) and torch.equal(g_base2, 2.0 * g_base)


@pytest.mark.parametrize("n", [4096 * 3 * 64 * 64, 1024 * 3 * 224 * 224 + 3])
def test_pgd_step_properties_at_sweep_size(n):
    gen = torch.Generator(device=DEV).manual_seed(9)
    eps, a = 16 / 255, 2 / 255
    x0 = torch.rand(n, device=DEV, generator=gen)
    x = torch.clamp(x0 + (torch.rand(n, device=DEV, generator=gen) * 2 - 1) * eps, 0, 1)
    g = torch.randn(n, device=DEV, generator=gen)
    g[::101] = 0
    y = F_ee.pgd_linf_step(x, g, x0, a, eps)
    ref = torch.clamp(torch.min(torch.max(x + a * torch.sign(g), x0 - eps), x0 + eps), 0, 1)   # attacks.py:25-27
    assert torch.equal(y, ref)
    assert float(y.min()) >= 0 and float(y.max()) <= 1
    assert float((y - x0).abs().max()) <= eps + 2.5e-7          # x0 +- eps is rounded to fp32 before the compare
    assert torch.equal(F_ee.pgd_linf_step(y, g, x0, 0.0, eps), y)               # projection is idempotent
    assert torch.equal(y[::101], x[::101])                                      # sign(0) = 0
    d = (x - x0).contiguous()
    adv = F_ee.free_at_step_(d, g, x0, a, eps)
    assert float(d.abs().max()) <= np.float32(eps) and torch.equal(adv, torch.clamp(x0 + d, 0, 1))


def test_cuda_graph_capture_of_one_pgd_iteration():
    """The library allocates nothing and never synchronises, so a whole hot-path iteration (fused forward,
    fused backward, PGD step) can be captured in a CUDA graph and replayed with identical results."""
    shape = (256, 3, 64, 64)                      # BASELINE configs[1] batch
    gen = torch.Generator(device=DEV).manual_seed(3)
    x0 = torch.rand(shape, device=DEV, generator=gen)
    x = torch.clamp(x0 + (torch.rand(shape, device=DEV, generator=gen) * 2 - 1) * (16 / 255), 0, 1)
    base = torch.rand(shape, device=DEV, generator=gen) * 1.1 - 0.1
    g_out = torch.randn(shape, device=DEV, generator=gen)
    p = F_ee.make_params("step125", O.gaussian3(), 0.0, None, T.HIGH, False)
    out, g_x, g_base, x_new = (torch.empty_like(x) for _ in range(4))

    def iteration():
        F_ee.edge_blend(x, base, p, 1.0, out=out)
        F_ee.edge_blend_backward(g_out, x, base, p, 1.0, g_x=g_x, g_base=g_base)
        F_ee.pgd_linf_step(x, g_x, x0, 2 / 255, 16 / 255, out=x_new)

    iteration()
    torch.cuda.synchronize()
    want = [t.clone() for t in (out, g_x, g_base, x_new)]
    for t in (out, g_x, g_base, x_new):
        t.zero_()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        iteration()                                # warm-up on the capture stream
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        iteration()
    for t in (out, g_x, g_base, x_new):
        t.zero_()
    graph.replay()
    torch.cuda.synchronize()
    for got, ref in zip((out, g_x, g_base, x_new), want):
        assert torch.equal(got, ref)


def test_thread_safety_and_current_stream():
    """DataParallel worker threads and autograd engine threads call into the library concurrently, each on
    its own current stream; the library must be re-entrant and enqueue on the caller's stream."""
    import threading
    shape = (64, 3, 64, 64)
    p = F_ee.make_params("canny", O.gaussian3(), 0.0, T.LOW, T.HIGH, True)
    results, errors = {}, []

    def worker(i):
        try:
            gen = torch.Generator(device=DEV).manual_seed(100 + i)
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                x = torch.rand(shape, device=DEV, generator=gen)
                base = torch.rand(shape, device=DEV, generator=gen)
                g = torch.randn(shape, device=DEV, generator=gen)
                outs = []
                for _ in range(20):
                    out = F_ee.edge_blend(x, base, p, 1.0)
                    gx, gb = F_ee.edge_blend_backward(g, x, base, p, 1.0)
                    outs.append((out, gx, gb))
                st.synchronize()
            for out, gx, gb in outs[1:]:
                assert torch.equal(out, outs[0][0]) and torch.equal(gx, outs[0][1]) and torch.equal(gb, outs[0][2])
            results[i] = (x.cpu().numpy(), base.cpu().numpy(), g.cpu().numpy(), outs[0][0].cpu().numpy(), outs[0][1].cpu().numpy())
        except Exception as e:      # pragma: no cover
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    po = O.make_params("canny", alpha=0.0, low=T.LOW, high=T.HIGH, hysteresis=True)
    for i, (x, base, g, out, gx) in results.items():
        assert np.array_equal(out, O.edge_blend_fwd(x, base, po, 1.0))
        assert np.array_equal(gx, O.edge_blend_bwd(g, x, base, po, 1.0)[0])


@pytest.mark.parametrize("variant", ["step125", "canny", "bpda"])
@pytest.mark.parametrize("shape,staging", [((512, 3, 64, 64), 0), ((512, 3, 64, 64), 6), ((256, 3, 32, 32), 0), ((512, 1, 28, 28), 0),
                                           ((24, 3, 224, 224), 0), ((24, 3, 224, 224), 7), ((24, 3, 224, 224), 5)], ids=str)
def test_repeated_launches_are_bit_identical(variant, shape, staging):
    """Poor man's race detector (compute-sanitizer is not available on the GPU pool): the kernels alias shared-memory
    regions between stages (S -> A, Bl -> GB, gx1 -> Bv, in-place channel sum after TMA staging) and read neighbour rows
    across barriers; a missing barrier or a region overlap shows up as run-to-run differences under load.  Ten launches on
    a batch large enough to fill the GPU several times over must agree bit for bit."""
    gen = torch.Generator(device=DEV).manual_seed(11)
    x = torch.rand(shape, device=DEV, generator=gen)
    base = torch.rand(shape, device=DEV, generator=gen) * 1.1 - 0.1
    g_out = torch.randn(shape, device=DEV, generator=gen)
    low = None if variant == "step125" else T.LOW
    p = F_ee.make_params(variant, O.gaussian3(), 0.0, low, T.HIGH, True)
    L = _lib.load()
    L.ee_set_tuning(0, 0, staging)
    ref = None
    for _ in range(10):
        out = F_ee.edge_blend(x, base, p, 1.0)
        g_x, g_base = F_ee.edge_blend_backward(g_out, x, base, p, 1.0)
        cur = (out, g_x, g_base)
        if ref is None:
            ref = cur
        else:
            for a, b in zip(ref, cur):
                assert torch.equal(a, b)
    L.ee_set_tuning(0, 0, 0)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_tensors_on_a_second_device_while_device_0_is_current():
    """DataParallel replicas and DDP ranks hand the library tensors of a device that is not the thread's current one: the
    wrappers switch device and stream per call (functional._on_device / _stream), per-device state (function attributes
    for > 48 KB of shared memory, the low-pass tables) must follow."""
    from edge_enhancement_b200 import core
    assert torch.cuda.current_device() == 0
    d1 = torch.device("cuda:1")
    x, base, g_out, _ = T.make_inputs(77, 6, 3, 64, 64)
    for variant in ("step125", "canny"):
        low = None if variant == "step125" else T.LOW
        p = F_ee.make_params(variant, O.gaussian3(), 0.0, low, T.HIGH, True)
        po = O.make_params(variant, alpha=0.0, low=low, high=T.HIGH, hysteresis=True)
        out = F_ee.edge_blend(torch.from_numpy(x).to(d1), torch.from_numpy(base).to(d1), p, 1.0)
        g_x, g_base = F_ee.edge_blend_backward(torch.from_numpy(g_out).to(d1), torch.from_numpy(x).to(d1), torch.from_numpy(base).to(d1), p, 1.0)
        assert out.device == d1 and np.array_equal(out.cpu().numpy(), O.edge_blend_fwd(x, base, po, 1.0))
        o_gx, o_gb = O.edge_blend_bwd(g_out, x, base, po, 1.0)
        assert np.array_equal(g_x.cpu().numpy(), o_gx) and np.array_equal(g_base.cpu().numpy(), o_gb)
    y = F_ee.hfs(torch.from_numpy(x).to(d1), 8)
    assert y.device == d1 and np.array_equal(y.cpu().numpy(), O.hfs(x, 8))
    xw = torch.rand(2, 3, 224, 224, device=d1)
    gw = torch.randn_like(xw)
    p = F_ee.make_params("step125", O.gaussian3(), 0.0, None, T.HIGH, True)
    gx1, _ = F_ee.edge_blend_backward(gw, xw, xw, p, 1.0)                       # TMA-staged tiles on device 1
    gx0, _ = F_ee.edge_blend_backward(gw.to("cuda:0"), xw.to("cuda:0"), xw.to("cuda:0"), p, 1.0)
    assert torch.equal(gx1.cpu(), gx0.cpu())
    assert torch.cuda.current_device() == 0


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_dataparallel_replicas_use_the_fused_front_end():
    """The reference wraps its MNIST / Tiny-ImageNet models in nn.DataParallel (experiments_tinyimagenet.py:110): the
    front end then runs on worker THREADS, one per device, each with its own current device and stream."""
    import contextlib, io
    from edge_enhancement_b200 import core

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            with contextlib.redirect_stdout(io.StringIO()):
                self.front = core.EdgeEnhance(cize=64, r=8, w=1.0, low=38.0, high=76.0, alpha=0.0, sigma=1,
                                              type_canny='CannyFilter_step125_1')
            self.conv = torch.nn.Conv2d(3, 4, 3, padding=1)

        def forward(self, x):
            return self.conv(self.front(x)).mean((1, 2, 3))

    torch.manual_seed(0)
    net = Net().to("cuda:0")
    x = torch.rand(16, 3, 64, 64, device="cuda:0")
    xa = x.clone().requires_grad_()
    net(xa).sum().backward()
    dp = torch.nn.DataParallel(net, device_ids=[0, 1])
    xb = x.clone().requires_grad_()
    dp(xb).sum().backward()
    assert torch.allclose(xa.grad, xb.grad, rtol=1e-5, atol=1e-7)
    with torch.no_grad():
        assert torch.allclose(net(x), dp(x), rtol=1e-5, atol=1e-6)
