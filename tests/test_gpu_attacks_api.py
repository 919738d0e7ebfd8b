"""-m gpu: every attack of the drop-in `attacks` module through its reference call signature, against
the reference's update lines written out with stock torch ops on the same device, same seed, same model.

The reference attack functions themselves cannot run on the GPU box (no /root/reference there; several
hard-code device='cuda' and cannot run on the CPU box either), so the comparison object is the literal
torch expression of each reference line (cited).  The updates are bit-identical; differences could only
come from the model's gradient, which is the same tensor on both sides here."""
import contextlib
import io
import types

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from edge_enhancement_b200 import attacks, core, functional as F_ee   # noqa: E402

DEV = "cuda:0"
EPS, STEP = 16 / 255, 2 / 255


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


class TinyEE(nn.Module):
    """edge-enhanced front end (fused drop-in) + a small conv net; deterministic init."""

    def __init__(self, n_class=10, variant="step125"):
        super().__init__()
        torch.manual_seed(0)
        cls = {"step125": core.CannyFilter_step125_1, "canny": core.CannyFilter}[variant]
        self.canny = quiet(cls, use_cuda=False)
        self.conv = nn.Conv2d(3, 8, 3, padding=1)
        self.fc = nn.Linear(8, n_class)

    def forward(self, x):
        z = core.edge_enhance(x, x, self.canny, 1.0, 38 / 255, 76 / 255, True)
        h = F.relu(self.conv(z)).mean((2, 3))
        return self.fc(h)


@pytest.fixture(scope="module")
def setup():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.deterministic = True
    model = TinyEE().to(DEV).eval()
    g = torch.Generator(device=DEV).manual_seed(5)
    x = torch.rand(6, 3, 32, 32, device=DEV, generator=g)
    y = torch.randint(0, 10, (6,), device=DEV, generator=g)
    return model, x, y


def ref_linf(x, g, x0, step, eps):
    x = x.detach() + step * torch.sign(g.detach())                   # utils/attacks.py:25
    x = torch.min(torch.max(x, x0 - eps), x0 + eps)                  # :26
    return torch.clamp(x, 0, 1)                                      # :27


def in_ball(x_adv, x, eps):
    return float(x_adv.min()) >= 0 and float(x_adv.max()) <= 1 and float((x_adv - x).abs().max()) <= eps + 3e-7


def test_pgd_random_start_matches_torch_expression(setup):
    model, x, y = setup
    args = types.SimpleNamespace(random=True, epsilon=EPS)
    torch.manual_seed(11)
    got = attacks.PGD(model, args, x, y, 5, STEP)
    torch.manual_seed(11)
    xr = x.detach() + torch.zeros_like(x).uniform_(-EPS, EPS)        # :15-17
    xr = torch.clamp(xr, 0, 1)
    for _ in range(5):
        xr.requires_grad_()
        loss = F.cross_entropy(model(xr), y, reduction='sum')
        g = torch.autograd.grad(loss, [xr])[0]
        xr = ref_linf(xr, g, x, STEP, EPS)
    assert torch.equal(got, xr) and in_ball(got, x, EPS)


def test_targeted_pgd_and_trick(setup):
    model, x, y = setup
    args = types.SimpleNamespace(random=True, epsilon=EPS, prob_start_from_clean=0.0)
    for fn in (attacks.targeted_PGD, attacks.targeted_PGD_trick):
        torch.manual_seed(3)
        adv, tgt = fn(model, args, x, y, 3, STEP, 10, DEV)
        assert tgt.shape == y.shape and bool((tgt != y).all()) and in_ball(adv, x, EPS)
    # descends: targeted loss after the attack is not larger than before
    torch.manual_seed(3)
    args.random = False
    adv, tgt = attacks.targeted_PGD(model, args, x, y, 5, STEP, 10, DEV)
    with torch.no_grad():
        assert F.cross_entropy(model(adv), tgt) <= F.cross_entropy(model(x), tgt) + 1e-6


def test_fgsm(setup):
    model, x, y = setup
    xr = x.detach().requires_grad_()
    g = torch.autograd.grad(F.cross_entropy(model(xr), y, reduction='sum'), [xr])[0]
    for targeted, sgn in ((False, 1.0), (True, -1.0)):
        got = attacks.FGSM(model, x, y, targeted=targeted, step_size=0.007)
        want = torch.clamp(x + sgn * 0.007 * torch.sign(g), 0.0, 1.0)                # :121-126
        assert torch.equal(got, want)


def test_trades_alp_classes(setup):
    model, x, y = setup
    with torch.no_grad():
        logits = model(x)
    tr = attacks.Trades(step_size=STEP, epsilon=EPS, perturb_steps=3, beta=6.0)
    torch.manual_seed(7)
    got = tr.PGD_Linf(model, x, logits)
    torch.manual_seed(7)
    xa = x.detach() + 0.001 * torch.randn(x.shape, device=x.device)                 # :406
    prob = F.softmax(logits, dim=-1)
    for _ in range(3):
        xa.requires_grad_()
        loss = nn.KLDivLoss(reduction="batchmean")(F.log_softmax(model(xa), dim=1), prob)      # :412
        g = torch.autograd.grad(loss, [xa])[0].detach()
        xa = ref_linf(xa, g, x, STEP, EPS)                                          # :414-416
    assert torch.equal(got, xa)
    torch.manual_seed(7)
    l2 = tr.PGD_L2(model, x, logits)
    assert l2.shape == x.shape and float(l2.min()) >= 0 and float(l2.max()) <= 1
    delta_rms = attacks.l2_norm(l2 - x)
    assert float(delta_rms.max()) <= EPS * (1 + 1e-4)                               # :394-397 projection
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    assert torch.isfinite(tr.loss(model, logits, got, y, opt))
    model.eval()
    for cls in (attacks.ALP, attacks.targeted_ALP):
        a = cls(step_size=STEP, epsilon=EPS, perturb_steps=2, beta=1.0)
        torch.manual_seed(1)
        adv = a.PGD_Linf(model, x, y)
        assert in_ball(adv, x, EPS + 0.005)
        a.reset_steps(1)
        assert a.perturb_steps == 1
    t = attacks.targeted_ALP(step_size=STEP, epsilon=EPS, perturb_steps=2, n_class=10)
    assert in_ball(t.tarPGD_Linf(model, x, y, DEV), x, EPS + 0.005)
    model1000 = TinyEE(n_class=1000).to(DEV).eval()            # tar_alp_imagenet hard-codes 1000 classes (:341-342)
    adv, tgt = attacks.tar_alp_imagenet(model1000, types.SimpleNamespace(epsilon=EPS), x, y, 2, STEP, DEV)
    assert in_ball(adv, x, EPS + 0.005) and tgt.shape == y.shape and int(tgt.max()) < 1000
    model.eval()


def test_avmixup(setup):
    model, x, y = setup
    args = types.SimpleNamespace(random=True, epsilon=EPS)
    av = attacks.AVmixup(args, gamma=2.0, lambda1=1.0, lambda2=0.1, step_size=STEP, num_steps=2, num_classes=10, device=DEV)
    onehot = F.one_hot(y, 10).float()
    np.random.seed(0)
    torch.manual_seed(2)
    xm, ym = av.perturb(model, x, onehot)
    assert xm.shape == x.shape and xm.dtype == torch.float32 and ym.shape == (6, 10)
    assert float(xm.min()) >= 0 and float(xm.max()) <= 1
    assert torch.allclose(ym.sum(1).float(), torch.ones(6, device=DEV), atol=1e-5)
    xm2, _ = av.tar_perturb(model, x, onehot)
    assert xm2.shape == x.shape


def test_cw_linf(setup):
    model, x, y = setup
    with torch.no_grad():
        pred = model(x).argmax(1)
    target = (pred + 1) % 10
    torch.manual_seed(4)
    adv, p = attacks.CWLinfAttack(x, pred, model, 0.03, None, 0.05, max_iters=3, target=target, n_class=10, cur_device=DEV)
    assert adv.shape == x.shape and p.shape == x.shape
    assert float(adv.min()) >= 0 and float(adv.max()) <= 1
    assert float((adv - x).abs().max()) <= 0.03 + 1e-6                                # :218 projection


def test_free_at_update_matches_reference_lines(setup):
    model, x, y = setup
    B = x.shape[0]
    fgsm_step, clip = 4 / 255, 4 / 255
    noise = (torch.rand(B + 2, 3, 32, 32, device=DEV) * 2 - 1) * clip                 # global buffer, AT_hfs_...py:286
    ref_noise = noise.clone()
    nb = noise[0:B].clone().requires_grad_()
    in1 = (x + nb).clamp(0, 1.0)                                                      # :314-315
    F.cross_entropy(model(in1), y).backward()                                         # :327
    ref_noise[0:B] += fgsm_step * torch.sign(nb.grad)                                 # :330-331
    ref_noise.clamp_(-clip, clip)                                                     # :332
    ref_in1 = (x + ref_noise[0:B]).clamp(0, 1.0)                                      # next repeat :314-315
    nxt = attacks.free_at_update_(noise, nb.grad, x, fgsm_step, clip)
    assert torch.equal(noise, ref_noise) and torch.equal(nxt, ref_in1)


def test_model_level_gradients_match_unfused_torch_path(setup):
    """d loss / d x through the fused edge_enhance equals the gradient through the module-level filter +
    torch blend (resnet_EE.py:182-191 written out), i.e. the fused backward is the adjoint of the same op."""
    model, x, y = setup
    x1 = x.detach().requires_grad_()
    g1 = torch.autograd.grad(F.cross_entropy(model(x1), y), [x1])[0]
    x2 = x.detach().requires_grad_()
    e = model.canny(x2, low_threshold=38 / 255, high_threshold=76 / 255, hysteresis=True)
    z = torch.clamp(x2 + 1.0 * e, 0.0, 1.0)
    h = F.relu(model.conv(z)).mean((2, 3))
    g2 = torch.autograd.grad(F.cross_entropy(model.fc(h), y), [x2])[0]
    assert torch.allclose(g1, g2, rtol=1e-5, atol=1e-7 * float(g2.abs().max()))


def test_graphed_pgd_matches_eager_pgd():
    """attacks.GraphedPGD (one captured iteration replayed num_steps times) returns exactly what attacks.PGD returns:
    same kernels, same order, only the launches come from a CUDA graph."""
    import contextlib, io
    B, C, S, n_class = 64, 3, 32, 10
    gen = torch.Generator(device=DEV).manual_seed(21)
    x = torch.rand((B, C, S, S), device=DEV, generator=gen)
    y = torch.randint(0, n_class, (B,), device=DEV, generator=gen)
    with contextlib.redirect_stdout(io.StringIO()):
        canny = core.CannyFilter_step125_1(use_cuda=False, alpha=0.0)
    weight = torch.randn((n_class, C * S * S), device=DEV, generator=gen) * 0.05

    def model(inp):
        z = core.edge_enhance(inp, inp, canny, 1.0, None, 76 / 255, True)
        return z.reshape(z.shape[0], -1) @ weight.t()

    class A:
        random = False
        epsilon = 16 / 255

    want = attacks.PGD(model, A, x, y, 7, 2 / 255)
    pgd = attacks.GraphedPGD(model, A, x, y, 2 / 255)
    got = pgd(x, y, 7)
    assert torch.equal(got, want)
    # a second batch through the same captured graph
    x2 = torch.rand((B, C, S, S), device=DEV, generator=gen)
    y2 = torch.randint(0, n_class, (B,), device=DEV, generator=gen)
    assert torch.equal(pgd(x2, y2, 5), attacks.PGD(model, A, x2, y2, 5, 2 / 255))


def test_graphed_pgd_with_the_trades_kl_loss():
    """GraphedPGD with loss_fn = the KL loss of Trades.PGD_Linf (attacks.py:409-416) and its start point reproduces
    the eager TRADES inner loop bit for bit."""
    B, C, S, n_class = 32, 3, 32, 10
    gen = torch.Generator(device=DEV).manual_seed(33)
    x = torch.rand((B, C, S, S), device=DEV, generator=gen)
    with contextlib.redirect_stdout(io.StringIO()):
        canny = core.CannyFilter_step125_1(use_cuda=False, alpha=0.0)
    weight = torch.randn((n_class, C * S * S), device=DEV, generator=gen) * 0.05

    def model(inp):
        z = core.edge_enhance(inp, inp, canny, 1.0, None, 76 / 255, True)
        return z.reshape(z.shape[0], -1) @ weight.t()

    eps, step, steps = 8 / 255, 2 / 255, 5
    with torch.no_grad():
        preds = model(x)
    x_init = x + 0.001 * torch.randn(x.shape, device=DEV, generator=gen)
    kl = nn.KLDivLoss(reduction='sum')

    # eager loop exactly as Trades.PGD_Linf writes it (with the fused update)
    xa = x_init.clone()
    for _ in range(steps):
        xa.requires_grad_()
        with torch.enable_grad():
            loss = kl(F.log_softmax(model(xa), dim=1), F.softmax(preds, dim=1))
        g = torch.autograd.grad(loss, [xa])[0]
        xa = F_ee.pgd_linf_step(xa.detach(), g.detach(), x.detach(), step, eps, 0.0, 1.0)

    args = types.SimpleNamespace(random=False, epsilon=eps)
    soft = F.softmax(preds, dim=1)
    pgd = attacks.GraphedPGD(model, args, x, soft, step, loss_fn=lambda logits, y: kl(F.log_softmax(logits, dim=1), y))
    assert torch.equal(pgd(x, soft, steps, x_init=x_init), xa)
