"""Generate tests/golden/*.npz by running the UNMODIFIED reference (utils/core.py, utils/attacks.py)
on CPU in the build container.

TEST INFRASTRUCTURE ONLY.  The reference cannot travel to the GPU box, so its outputs are committed
as small fixtures together with this script.  Re-run with:  python -m oracle.make_golden

Each edge case stores inputs (x, base, g_out) and the reference's outputs:
    edge   = Filter(x, low, high, hysteresis)                       utils/core.py forward
    out    = clamp(base + w*edge, 0, 1)                             resnet_EE.py:189-191
    g_x    = d<out, g_out>/dx  (edge path only, base held constant) autograd
    g_base = d<out, g_out>/dbase
The attack cases store the torch expressions of utils/attacks.py evaluated on CPU, and one full
reference PGD run against a tiny fixed model whose front end is the reference filter + blend.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from tests import common as T  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# name, class name, variant, (B,C,H,W), input kind, alpha, sigma, low, high, hysteresis, w, seed
EDGE_CASES = [
    ("step125_tiny_3x32", "CannyFilter_step125_1", "step125", (4, 3, 32, 32), "uniform", 0.0, 1.0, None, 76 / 255, False, 1.0, 101),
    ("step125_tiny_3x64", "CannyFilter_step125_1", "step125", (2, 3, 64, 64), "uniform", 0.0, 1.0, None, 76 / 255, False, 1.0, 102),
    ("step125_mnist_1x28", "CannyFilter_step125_1", "step125", (8, 1, 28, 28), "sparse", 0.3, 1.0, None, 51 / 255, False, 1.0, 103),
    ("step125_smooth_w05", "CannyFilter_step125_1", "step125", (2, 3, 48, 40), "smooth", 0.0, 1.0, None, 76 / 255, False, 0.5, 104),
    ("step125_odd_17x23", "CannyFilter_step125_1", "step125", (2, 3, 17, 23), "uniform", 0.05, 1.0, None, 60 / 255, False, 1.0, 105),
    ("canny_tiny_3x32", "CannyFilter", "canny", (4, 3, 32, 32), "uniform", 0.0, 1.0, 38 / 255, 76 / 255, True, 1.0, 201),
    ("canny_tiny_3x64", "CannyFilter", "canny", (2, 3, 64, 64), "uniform", 0.0, 1.0, 38 / 255, 76 / 255, True, 1.0, 202),
    ("canny_mnist_1x28", "CannyFilter", "canny", (8, 1, 28, 28), "sparse", 0.3, 1.0, 25 / 255, 51 / 255, True, 1.0, 203),
    ("canny_smooth", "CannyFilter", "canny", (2, 3, 48, 40), "smooth", 0.0, 1.0, 38 / 255, 76 / 255, True, 1.0, 204),
    ("canny_mix", "CannyFilter", "canny", (2, 3, 24, 24), "uniform", 0.0, 1.0, 38 / 255, 76 / 255, False, 1.0, 205),
    ("canny_low", "CannyFilter", "canny", (2, 3, 24, 24), "uniform", 0.0, 1.0, 38 / 255, None, False, 1.0, 206),
    ("canny_raw", "CannyFilter", "canny", (2, 3, 24, 24), "uniform", 0.0, 1.0, None, None, False, 1.0, 207),
    ("canny_sigma2_odd", "CannyFilter", "canny", (2, 2, 19, 21), "uniform", 0.02, 2.0, 30 / 255, 60 / 255, True, 0.7, 208),
    ("bpda_tiny_3x32", "CannyFilter_BPDA", "bpda", (4, 3, 32, 32), "uniform", 0.0, 1.0, 38 / 255, 76 / 255, True, 1.0, 301),
    ("bpda_tiny_3x64", "CannyFilter_BPDA", "bpda", (2, 3, 64, 64), "uniform", 0.0, 1.0, 38 / 255, 76 / 255, True, 1.0, 302),
    ("bpda_smooth", "CannyFilter_BPDA", "bpda", (2, 3, 48, 40), "smooth", 0.0, 1.0, 38 / 255, 76 / 255, True, 1.0, 303),
    ("bpda_mix", "CannyFilter_BPDA", "bpda", (2, 3, 24, 24), "uniform", 0.0, 1.0, 38 / 255, 76 / 255, False, 1.0, 304),
    ("bpda_low_is_raw", "CannyFilter_BPDA", "bpda", (2, 3, 24, 24), "uniform", 0.0, 1.0, 38 / 255, None, False, 1.0, 305),
]


def run_edge_case(rc, case):
    name, cls, variant, shape, kind, alpha, sigma, low, high, hyst, w, seed = case
    x, base, g_out, _ = T.make_inputs(seed, *shape, kind=kind)
    with ref_loader.quiet():
        f = getattr(rc, cls)(sigma=sigma, use_cuda=False, alpha=alpha)
    xt = torch.from_numpy(x).requires_grad_()
    bt = torch.from_numpy(base).requires_grad_()
    edge = f(xt, low_threshold=low, high_threshold=high, hysteresis=hyst)
    out = torch.clamp(bt + w * edge, 0.0, 1.0)
    out.backward(torch.from_numpy(g_out))
    np.savez_compressed(
        os.path.join(OUT, "edge_%s.npz" % name),
        x=x, base=base, g_out=g_out,
        edge=edge.detach().numpy().astype(np.float32), out=out.detach().numpy(),
        g_x=xt.grad.numpy(), g_base=bt.grad.numpy(),
        meta=np.array([variant, str(alpha), str(sigma), repr(low), repr(high), str(int(hyst)), str(w)]))
    e = edge.detach().numpy()
    print("%-22s edge mean %.3f  finite g_x %.3f" % (name, e.mean(), np.isfinite(xt.grad.numpy()).mean()))


class TinyEENet(torch.nn.Module):
    """Reference filter + blend front end (resnet_EE.py:182-191 with base = x, no HFS) feeding a fixed
    linear head; weights come from a seeded numpy RNG so the GPU test can rebuild the same model."""

    def __init__(self, canny, low, high, w, C, H, W, n_class, seed):
        super().__init__()
        self.canny, self.low, self.high, self.w = canny, low, high, w
        r = np.random.default_rng(seed)
        self.weight = torch.from_numpy(r.standard_normal((n_class, C * H * W)).astype(np.float32) * 0.05)

    def forward(self, x):
        e = self.canny(x, low_threshold=self.low, high_threshold=self.high, hysteresis=True)
        z = torch.clamp(x + self.w * e, 0.0, 1.0)
        return z.reshape(z.shape[0], -1) @ self.weight.to(z.device).t()


def run_attack_cases(rc, ra):
    eps, a = 16 / 255, 2 / 255
    x, g, x0 = T.make_attack_inputs(401, (4, 3, 16, 16), eps)
    g.reshape(-1)[7] = np.nan
    tx, tg, tx0 = map(torch.from_numpy, (x, g, x0))
    pgd = torch.clamp(torch.min(torch.max(tx + a * torch.sign(tg), tx0 - eps), tx0 + eps), 0, 1)     # attacks.py:25-27
    tpgd = torch.clamp(torch.min(torch.max(tx - a * torch.sign(tg), tx0 - eps), tx0 + eps), 0.0, 1.0)  # :52-54
    fgsm = torch.clamp(tx + 0.007 * torch.sign(tg), 0.0, 1.0)                                          # :124-126
    delta = (tx - tx0).clone()
    delta += (4 / 255) * torch.sign(tg)                      # AT_hfs_canny_free_imagenet_ddp.py:330-331
    delta.clamp_(-4 / 255, 4 / 255)                          # :332
    free_adv = (tx0 + delta).clamp_(0, 1.0)                  # :314-315
    # TRADES L2, attacks.py:391-399 (g without the NaN)
    g2 = g.copy(); g2.reshape(-1)[7] = 0.25
    tg2 = torch.from_numpy(g2)
    gr = tg2 / (ra.l2_norm(tg2).unsqueeze(-1).unsqueeze(-1).unsqueeze(-1) + 1e-8)
    xa = tx + 0.5 * gr
    d = xa - tx0
    dn = ra.l2_norm(d)
    cond = dn > 0.02
    d[cond] *= 0.02 / dn[cond].unsqueeze(-1).unsqueeze(-1).unsqueeze(-1)
    l2 = torch.clamp(tx0 + d, 0.0, 1.0)
    # CW, attacks.py:213-222
    mx, mn = tx0 + 0.03, tx0 - 0.03
    cw = tx + 0.00392 * torch.sign(tg)
    cw = torch.max(torch.min(cw, tx0 + 0.02), tx0 - 0.02)
    cw = cw.clamp(0, 1)
    cw = torch.max(torch.min(cw, mx), mn)
    np.savez_compressed(os.path.join(OUT, "attack_steps.npz"), x=x, g=g, x0=x0, g2=g2, pgd=pgd.numpy(),
                        tpgd=tpgd.numpy(), fgsm=fgsm.numpy(), free_delta=delta.numpy(), free_adv=free_adv.numpy(),
                        l2=l2.numpy(), cw=cw.numpy(), cw_min=mn.numpy(), cw_max=mx.numpy())

    # full reference PGD-10 through the reference filter (deterministic start: args.random = False)
    class Args:
        random = False
        epsilon = 16 / 255
    B, C, H, W, n_class = 4, 3, 16, 16, 10
    x = T.rng(402).random((B, C, H, W), dtype=np.float32)
    y = T.rng(403).integers(0, n_class, size=(B,))
    for variant, cls in (("step125", "CannyFilter_step125_1"), ("canny", "CannyFilter")):
        with ref_loader.quiet():
            canny = getattr(rc, cls)(use_cuda=False, alpha=0.0)
        model = TinyEENet(canny, 38 / 255, 76 / 255, 1.0, C, H, W, n_class, seed=404)
        xadv = ra.PGD(model, Args, torch.from_numpy(x), torch.from_numpy(y), 10, 2 / 255)
        # gradient magnitudes along the way are needed to identify sign-ambiguous elements
        np.savez_compressed(os.path.join(OUT, "pgd10_%s.npz" % variant), x=x, y=y, x_adv=xadv.numpy(),
                            head_seed=np.array(404), n_class=np.array(n_class))
        print("pgd10_%s: mean |x_adv - x| = %.4f" % (variant, (xadv - torch.from_numpy(x)).abs().mean()))


def run_attack_extras():
    """The torch expressions of the attack-loop host side, evaluated on CPU exactly as the
    reference writes them: random start (attacks.py:15-17) and the AVmixup vertex / mix (attacks.py:469-478)."""
    eps, gamma = 16 / 255, 2.0
    x, _, x0 = T.make_attack_inputs(601, (6, 3, 8, 8), eps)
    tx, tx0 = torch.from_numpy(x), torch.from_numpy(x0)
    torch.manual_seed(601)
    noise = torch.zeros_like(tx0).uniform_(-eps, eps)
    start = torch.clamp(tx0 + noise, 0, 1)
    x_weight = np.random.default_rng(602).beta(1.0, 1.0, [x.shape[0], 1, 1, 1])
    perturb = (tx - tx0) * gamma
    vertex = torch.clamp(tx0 + perturb, 0, 1)
    w = torch.from_numpy(x_weight)
    mixed = (tx0 * w + vertex * (1 - w)).to(torch.float)
    np.savez_compressed(os.path.join(OUT, "attack_extras.npz"), x=x, x0=x0, noise=noise.numpy(), start=start.numpy(),
                        weight=x_weight, mixed=mixed.numpy(), gamma=np.array(gamma), eps=np.array(eps))


def run_add_square_cases(rc):
    """Add_Square (utils/core.py:589-655) hard-codes .cuda(); on this CPU-only container Tensor.cuda is patched to
    the identity for the duration of the call, nothing else of the reference is touched.  The fixture stores the
    generator seed, so the drop-in (which draws with the same calls in the same order) must reproduce `out` from
    the seed alone, plus the draws themselves (stripe, table) for the kernel-level tests."""
    import math
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        for name, (B, C, S), eps, nq, seed in (("tiny_q1", (4, 3, 64), 16 / 255, 1, 501),
                                               ("mnist_q1", (8, 1, 28), 0.3, 1, 502),
                                               ("small_q5", (3, 3, 16), 8 / 255, 5, 503)):
            m = rc.Add_Square(channels=C, size=S, epsilon=eps, n_queries=nq)
            r = T.rng(seed)
            x = r.random((B, C, S, S), dtype=np.float32)
            x.reshape(-1)[::7] = 0.0                     # exercise both clamp sides
            x.reshape(-1)[3::11] = 1.0
            g = r.standard_normal(x.shape, dtype=np.float32)
            xt = torch.from_numpy(x).requires_grad_()
            torch.manual_seed(seed)
            out = m(xt)
            out.backward(torch.from_numpy(g))
            # replay the draws (same generator state) to store them
            torch.manual_seed(seed)
            stripe = torch.sign(2 * torch.rand([B, C, 1, S]) - 1).reshape(B, C, S).numpy()
            table = np.zeros((nq, 2 + C), np.float32)
            for it in range(nq):
                p = m.p_selection(it)
                s = max(int(round(math.sqrt(p * (C * S * S) / C))), 1)
                vh = int((0 + (S - s) * torch.rand([1])).long())
                sg = torch.sign(2 * torch.rand([C, 1, 1]) - 1).reshape(-1).numpy()
                table[it, 0], table[it, 1], table[it, 2:] = vh, s, np.float32(2. * eps) * sg
            np.savez_compressed(os.path.join(OUT, "add_square_%s.npz" % name), x=x, g=g, out=out.detach().numpy(),
                                g_x=xt.grad.numpy(), stripe=stripe.astype(np.float32), table=table,
                                meta=np.array([str(C), str(S), repr(eps), str(nq), str(seed)]))
            print("add_square_%s: mean |out - x| = %.4f" % (name, float((out.detach() - xt.detach()).abs().mean())))
    finally:
        torch.Tensor.cuda = orig_cuda


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(1)        # fixed reduction order for the fixtures
    rc, ra = ref_loader.load()
    for case in EDGE_CASES:
        run_edge_case(rc, case)
    run_attack_cases(rc, ra)
    run_add_square_cases(rc)
    run_attack_extras()
    print("fixtures written to", OUT)


if __name__ == "__main__":
    main()
