"""Drop-in for the reference's ``utils/attacks.py`` -- same names and call signatures.

The reference spells the same L-inf loop out eleven times (PGD :19-27, targeted_PGD :48-54, targeted_PGD_trick :78-84,
ALP :252-259, targeted_ALP :293-300 / :313-320, tar_alp_imagenet :348-355, Trades :408-416, AVmixup :458-468 / :497-507)
and its update lines cost 6-14 eager kernels per iteration.  Here every attack is a thin description -- start point,
loss, sign of the step -- handed to ONE loop (`_linf_attack`) whose update is ONE fused CUDA kernel
(libedge_b200.so: ee_pgd_linf_step_f32).  The model forward / backward in between stays on stock PyTorch: the CNN is
outside the product.  Random numbers are drawn with the same torch / numpy calls in the same order as the reference, so
a seeded run follows the same trajectory.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as F_ee


# --------------------------------------------------------------------------------------------------------------------
# the pieces every attack is assembled from
# --------------------------------------------------------------------------------------------------------------------
def _ce_sum(labels):
    return lambda logits: F.cross_entropy(logits, labels, reduction='sum')


def _ce_mean(labels):
    return lambda logits: F.cross_entropy(logits, labels)


def _linf_attack(model, loss_of_logits, inputs, x, num_steps, step_signed, epsilon):
    """num_steps times: g = d loss(model(x)) / dx ; x = clamp(min(max(x + step_signed*sign(g), inputs-eps), inputs+eps), 0, 1).
    A positive step ascends the loss (untargeted), a negative one descends towards the target labels inside the loss."""
    anchor = inputs.detach()
    for _ in range(num_steps):
        x.requires_grad_()
        with torch.enable_grad():
            loss = loss_of_logits(model(x))
        grad = torch.autograd.grad(loss, [x])[0]
        x = F_ee.pgd_linf_step(x.detach(), grad.detach(), anchor, step_signed, epsilon, 0.0, 1.0)
    return x


def _random_start(x, epsilon):
    """x + U(-eps, eps) clamped to [0, 1] (utils/attacks.py:15-17).  The noise is the reference's own draw
    (zeros_like(x).uniform_ advances the generator identically); add + clamp are one kernel."""
    return F_ee.add_clamp(x, torch.empty_like(x).uniform_(-epsilon, epsilon), 0.0, 1.0)


def _small_randn_start(x_natural):
    """x + 0.001 * randn (ALP / TRADES).  The reference hard-codes device='cuda' (attacks.py:250, :291, :311, :383, :406);
    the draw follows the input's device instead."""
    return x_natural.detach() + 0.001 * torch.randn(x_natural.shape, device=x_natural.device).detach()


def _random_targets(labels, n_class, device):
    """A uniformly random label different from the true one (CPU randint, like the reference: attacks.py:35-36)."""
    offset = torch.randint(low=1, high=n_class, size=labels.shape).to(device)
    return torch.fmod(labels + offset, n_class)


def _start(args, inputs):
    x = inputs.detach()
    return _random_start(x, args.epsilon) if args.random else x


# --------------------------------------------------------------------------------------------------------------------
# utils/attacks.py:12-86 -- PGD and its targeted variants
# --------------------------------------------------------------------------------------------------------------------
def PGD(model, args, inputs, targets, num_steps, step_size):
    return _linf_attack(model, _ce_sum(targets), inputs, _start(args, inputs), num_steps, step_size, args.epsilon)


def targeted_PGD(model, args, inputs, labels, num_steps, step_size, nclass, device):
    target_labels = _random_targets(labels, nclass, device)          # drawn before the random start, like :35-47
    x = _linf_attack(model, _ce_sum(target_labels), inputs, _start(args, inputs), num_steps, -step_size, args.epsilon)
    return x, target_labels


def targeted_PGD_trick(model, args, inputs, labels, num_steps, step_size, nclass, device):
    """As targeted_PGD, but the random start is skipped with probability args.prob_start_from_clean (:74-77; this start
    keeps the reference's own draws -- torch.Tensor(shape).uniform_ on the CPU, then one torch.rand([]))."""
    target_labels = _random_targets(labels, nclass, device)
    x = inputs.detach()
    if args.random:
        noise = torch.Tensor(x.shape).uniform_(-args.epsilon, args.epsilon).to(device)
        from_noise = torch.gt(torch.rand([]), args.prob_start_from_clean).type(torch.float32).to(device)
        x = torch.clamp(x + from_noise * noise, 0.0, 1.0)
    x = _linf_attack(model, _ce_sum(target_labels), inputs, x, num_steps, -step_size, args.epsilon)
    return x, target_labels


def tar_alp_imagenet(model, args, inputs, labels, num_steps, step_size, device):
    """utils/attacks.py:337-357: targeted, 1000 classes, start x + 0.001 * randn (a CPU draw moved to `device`)."""
    target_labels = _random_targets(labels, 1000, device)
    x = inputs.detach() + 0.001 * torch.randn(inputs.shape).to(device).detach()
    x = _linf_attack(model, _ce_sum(target_labels), inputs, x, num_steps, -step_size, args.epsilon)
    return x, target_labels


class GraphedPGD:
    """CUDA-graph replay of the PGD loop of utils/attacks.py:19-27 for launch-bound batch sizes.

    One iteration -- model forward, cross-entropy (reduction='sum'), input gradient, fused sign/project/clamp update -- is
    captured once on static buffers and replayed `num_steps` times: at the reference's batch sizes (MNIST 128 x 1 x 28 x 28,
    Tiny-ImageNet 256 x 3 x 64 x 64) an iteration is a few hundred short kernels whose launch overhead, not their run time,
    sets the pace.  libedge_b200.so allocates nothing and never synchronises, so it is capturable as is.  The model's
    parameters are read in place (weight updates between calls are seen by the replays); the model must not change its
    control flow or allocate persistent state in forward.  Same arithmetic, same order as `PGD`.

        pgd = GraphedPGD(model, args, inputs, targets, step_size)      # captures
        x_adv = pgd(inputs, targets, num_steps)                        # replays

    `loss_fn(logits, targets)` replaces the cross-entropy (e.g. the KL loss of Trades.PGD_Linf, attacks.py:412, with
    `targets` = the clean softmax), a negative `step_size` gives the targeted variants (:52, :505), and `x_init` in the call
    overrides the start point (ALP / TRADES start from x + 0.001 * randn, :254, :406).
    """

    def __init__(self, model, args, example_inputs, example_targets, step_size, warmup=3, loss_fn=None):
        self.model, self.args, self.step_size = model, args, step_size
        self.loss_fn = loss_fn if loss_fn is not None else (lambda logits, y: F.cross_entropy(logits, y, reduction='sum'))
        self.x0 = example_inputs.detach().clone()
        self.x = self.x0.clone()
        self.y = example_targets.detach().clone()
        side = torch.cuda.Stream(device=self.x.device)
        side.wait_stream(torch.cuda.current_stream(self.x.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._iteration()
        torch.cuda.current_stream(self.x.device).wait_stream(side)
        self.x.copy_(self.x0)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._iteration()
        self.x.copy_(self.x0)

    def _iteration(self):
        x = self.x.detach().requires_grad_()
        with torch.enable_grad():
            loss = self.loss_fn(self.model(x), self.y)
        grad = torch.autograd.grad(loss, [x])[0]
        F_ee.pgd_linf_step(self.x, grad, self.x0, self.step_size, self.args.epsilon, 0.0, 1.0, out=self.x)

    def __call__(self, inputs, targets, num_steps, x_init=None):
        self.x0.copy_(inputs.detach())
        self.y.copy_(targets)
        if x_init is not None:
            self.x.copy_(x_init.detach())
        elif self.args.random:
            self.x.copy_(_random_start(self.x0, self.args.epsilon))
        else:
            self.x.copy_(self.x0)
        for _ in range(num_steps):
            self.graph.replay()
        return self.x.clone()


# --------------------------------------------------------------------------------------------------------------------
# utils/attacks.py:89-128 -- label smoothing, FGSM
# --------------------------------------------------------------------------------------------------------------------
class LabelSmoothLoss(torch.nn.Module):
    """Cross-entropy against (1 - smoothing) on the label and smoothing / (n - 1) elsewhere (:89-101)."""

    def __init__(self, smoothing=0.0):
        super(LabelSmoothLoss, self).__init__()
        self.smoothing = smoothing

    def forward(self, input, target):
        # the off-label weight is evaluated in the tensor's dtype, (1 * smoothing) / (n - 1), like the reference (:97)
        soft = input.new_ones(input.size()) * self.smoothing / (input.size(-1) - 1.)
        soft.scatter_(-1, target.unsqueeze(-1), 1. - self.smoothing)
        return (-soft * F.log_softmax(input, dim=-1)).sum(dim=-1).mean()


def compute_loss_and_error(logits, label, label_smoothing=0.):
    return LabelSmoothLoss(label_smoothing)(logits, label.long())


def FGSM(model, inputs, target, targeted=False, step_size=0.007):
    """One signed step and the [0, 1] clamp, no eps projection (:110-128)."""
    x = inputs.detach().requires_grad_()
    with torch.enable_grad():
        loss = F.cross_entropy(model(x), target, reduction='sum')
    grad = torch.autograd.grad(loss, [x])[0]
    return F_ee.fgsm_step(x.detach(), grad.detach(), -step_size if targeted else step_size, 0.0, 1.0)


def predict_from_logits(logits, dim=1):
    return logits.max(dim=dim, keepdim=False)[1]


# --------------------------------------------------------------------------------------------------------------------
# utils/attacks.py:136-232 -- CW with L-inf norm (evaluation only)
# --------------------------------------------------------------------------------------------------------------------
def _as_batch(t, dims):
    """`t[index]` with a 0-dim index drops the batch axis (one surviving sample): put it back (:150-152, :157)."""
    return t if t.dim() == dims else t.unsqueeze(0)


def _cw_margin_loss(outputs, one_hot_y, one_hot_target):
    """-(sum relu(correct - wrong + 50)) with `wrong` = the target logit, or the best other logit (:190-203)."""
    correct = torch.sum(one_hot_y * outputs, dim=1)
    if one_hot_target is not None:
        wrong = torch.sum(one_hot_target * outputs, dim=1)
    else:
        wrong, _ = torch.max((1 - one_hot_y) * outputs - 1e4 * one_hot_y, dim=1)
    return -torch.sum(F.relu(correct - wrong + 50))


def _one_hot(labels, n_class, device):
    oh = torch.zeros(labels.size(0), n_class).to(device)
    oh[torch.arange(labels.size(0)), labels] = 1
    return oh


def CWLinfAttack(x, y, model, magnitude, previous_p, max_eps, max_iters=20, target=None, _type='linf',
                 n_class=10, cur_device=None):
    """Attacks only the samples the model still classifies correctly; the others are returned untouched.  Per iteration
    the reference's four projection / clamp lines (:212-222, hard-coded step 0.00392) are ONE fused kernel.  Returns
    (adv, perturbation); with `previous_p` the perturbation accumulates and the second projection is around x - previous_p."""
    model.eval()
    device = cur_device
    x, y = x.to(device), y.to(device)
    if target is not None:
        target = target.to(device)
    adv = x.clone()
    alive = predict_from_logits(model(x)) == y
    if torch.sum(alive).item() == 0:
        return adv, previous_p
    idx = alive.nonzero().squeeze()
    x, y = _as_batch(x[idx], 4), _as_batch(y[idx], 1)
    target = _as_batch(target[idx], 1)                       # like the reference, `target` is required in practice (:149)
    carried = None
    if previous_p is not None:
        carried = previous_p.to(device).clone()
        previous_p = _as_batch(carried[idx], 4)

    one_hot_y = _one_hot(y, n_class, device)
    one_hot_target = _one_hot(target, n_class, device) if target is not None else None
    mag = magnitude.item() if isinstance(magnitude, torch.Tensor) else magnitude
    x_d = x.detach()
    adv_imgs = (x_d + torch.FloatTensor(x.shape).uniform_(-mag, mag).to(device)).clamp_(0, 1)       # random start (:166-176)
    centre = x_d - previous_p if previous_p is not None else x_d
    min_x, max_x = (centre - max_eps).detach(), (centre + max_eps).detach()

    for _iter in range(int(max_iters)):
        adv_imgs.requires_grad_()
        with torch.enable_grad():
            loss = _cw_margin_loss(model(adv_imgs), one_hot_y, one_hot_target)
        grads = torch.autograd.grad(loss, adv_imgs, grad_outputs=None, only_inputs=True)[0]
        adv_imgs = F_ee.cw_linf_step(adv_imgs.detach(), grads.detach(), x_d, min_x, max_x, 0.00392, mag)
    adv_imgs = adv_imgs.clamp(0, 1)

    now_p = adv_imgs - x_d
    adv[idx] = adv_imgs
    if carried is not None:
        carried[idx] = previous_p + now_p
        return adv, carried
    return adv, now_p


# --------------------------------------------------------------------------------------------------------------------
# utils/attacks.py:236-333 -- ALP, targeted ALP
# --------------------------------------------------------------------------------------------------------------------
class ALP:
    def __init__(self, step_size=0.003, epsilon=0.047, perturb_steps=5, beta=1.0):
        self.step_size = step_size
        self.epsilon = epsilon
        self.perturb_steps = perturb_steps
        self.beta = beta

    def reset_steps(self, k):
        self.perturb_steps = k

    def _attack(self, model, x_natural, labels, step_signed):
        model.eval()
        return _linf_attack(model, _ce_mean(labels), x_natural, _small_randn_start(x_natural), self.perturb_steps,
                            step_signed, self.epsilon)

    def PGD_Linf(self, model, x_natural, y):
        return self._attack(model, x_natural, y, self.step_size)

    def loss(self, model, logits, logits_adv, y, optimizer):
        """0.5 CE(clean) + 0.5 CE(adv) + beta * MSE(logit pairing) (:263-272)."""
        model.train()
        optimizer.zero_grad()
        loss_robust = 0.5 * F.cross_entropy(logits, y) + 0.5 * F.cross_entropy(logits_adv, y)
        return loss_robust + self.beta * F.mse_loss(logits, logits_adv)


class targeted_ALP(ALP):
    def __init__(self, step_size=0.003, epsilon=0.047, perturb_steps=5, beta=1.0, n_class=200):
        super().__init__(step_size, epsilon, perturb_steps, beta)
        self.n_class = n_class

    def tarPGD_Linf(self, model, x_natural, y, device):
        model.eval()
        target_labels = _random_targets(y, self.n_class, device)        # drawn before the randn start (:307-311)
        return self._attack(model, x_natural, target_labels, -self.step_size)


# --------------------------------------------------------------------------------------------------------------------
# utils/attacks.py:360-429 -- TRADES
# --------------------------------------------------------------------------------------------------------------------
def squared_l2_norm(x):
    return (x.view(x.shape[0], -1) ** 2).mean(1)          # mean, not sum (:360-362)


def l2_norm(x):
    return squared_l2_norm(x).sqrt()


class Trades:
    def __init__(self, step_size=0.003, epsilon=0.047, perturb_steps=5, beta=1.0):
        self.step_size = step_size
        self.epsilon = epsilon
        self.perturb_steps = perturb_steps
        self.beta = beta
        self.criterion_kl = nn.KLDivLoss(reduction="batchmean")

    def reset_steps(self, k):
        self.perturb_steps = k

    def _kl_to(self, logits):
        prob = F.softmax(logits, dim=-1)
        return lambda adv_logits: self.criterion_kl(F.log_softmax(adv_logits, dim=1), prob)

    def PGD_Linf(self, model, x_natural, logits):
        model.eval()
        return _linf_attack(model, self._kl_to(logits), x_natural, _small_randn_start(x_natural), self.perturb_steps,
                            self.step_size, self.epsilon)

    def PGD_L2(self, model, x_natural, logits):
        """Normalised-gradient step with L2 re-projection (:381-401): per-sample RMS norms, step, projection and clamp are
        ONE kernel (ee_pgd_l2_step_f32: the sample stays on chip between the two norms)."""
        model.eval()
        x_adv = _small_randn_start(x_natural)
        kl = self._kl_to(logits)
        anchor = x_natural.detach()
        for _ in range(self.perturb_steps):
            x_adv.requires_grad_()
            with torch.enable_grad():
                loss = kl(model(x_adv))
            grad = torch.autograd.grad(loss, [x_adv])[0].detach()
            x_adv = F_ee.pgd_l2_step(x_adv.detach(), grad, anchor, self.step_size, self.epsilon)
        return x_adv

    def loss(self, model, logits, x_adv, labels, optimizer):
        model.train()
        optimizer.zero_grad()
        return F.cross_entropy(logits, labels) + self.beta * self._kl_to(logits)(model(x_adv))


# --------------------------------------------------------------------------------------------------------------------
# utils/attacks.py:433-518 -- AVmixup
# --------------------------------------------------------------------------------------------------------------------
class AVmixup:
    def __init__(self, args, gamma, lambda1, lambda2, step_size, num_steps, num_classes=200, device='cuda'):
        self.args = args
        self.gamma = gamma
        self.lambda1 = lambda1
        self.lambda2 = lambda2
        self.step_size = step_size
        self.num_steps = num_steps
        self.num_classes = num_classes
        self.device = device

    def _label_smoothing(self, one_hot, factor):
        return one_hot * factor + (one_hot - 1.) * ((factor - 1) / float(self.num_classes - 1))

    def _attack(self, model, inputs, soft_targets, step_signed):
        soft_ce = lambda logits: -torch.sum(F.log_softmax(logits, dim=1) * soft_targets)          # :462-463
        return _linf_attack(model, soft_ce, inputs, _start(self.args, inputs), self.num_steps, step_signed,
                            self.args.epsilon)

    def _mix(self, x, inputs, targets):
        """Adversarial vertex, per-sample Beta(1, 1) mix of inputs and labels (:469-478).  vertex, clamp, the float64 mix
        and the cast back are one kernel; the weights are the reference's numpy draw."""
        x_weight = np.random.beta(1.0, 1.0, [x.shape[0], 1, 1, 1])
        y_weight = torch.from_numpy(np.reshape(x_weight, [-1, 1])).to(x.device)
        x = F_ee.avmixup_mix(x.detach(), inputs.detach(), torch.from_numpy(x_weight).to(x.device), self.gamma)
        y = self._label_smoothing(targets, self.lambda1) * y_weight + self._label_smoothing(targets, self.lambda2) * (1 - y_weight)
        return x, y

    def perturb(self, model, inputs, targets):
        """:447-479 (targets are soft / one-hot labels)."""
        return self._mix(self._attack(model, inputs, targets, self.step_size), inputs, targets)

    def tar_perturb(self, model, inputs, targets):
        """:481-518: descends towards random target labels."""
        target_labels = _random_targets(targets, self.num_classes, self.device)
        return self._mix(self._attack(model, inputs, target_labels, -self.step_size), inputs, targets)


# --------------------------------------------------------------------------------------------------------------------
# free / fast adversarial training noise update (in-script loops of the reference):
#   ImageNet/free_imagenet/AT_hfs_canny_free_imagenet_ddp.py:312-315, :330-332
#   ImageNet/fgsm_imagenet/main_fast.py:233-235, :246-253 ; lib/utils.py:36-37
# --------------------------------------------------------------------------------------------------------------------
def free_at_update_(global_noise, noise_grad, inputs, fgsm_step, clip_eps):
    """global_noise[0:B] += fgsm_step*sign(noise_grad); clamp to +-clip_eps (in place) and return the
    next repeat's input clamp(inputs + global_noise[0:B], 0, 1), all in one kernel."""
    B = inputs.size(0)
    view = global_noise[0:B]
    return F_ee.free_at_step_(view, noise_grad.detach(), inputs.detach(), fgsm_step, clip_eps, 0.0, 1.0, True)
