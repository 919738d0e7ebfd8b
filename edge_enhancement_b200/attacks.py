"""Drop-in for the reference's ``utils/attacks.py`` -- same names and call signatures.

Every attack keeps the reference's control flow (random start, model forward, loss,
``torch.autograd.grad``) on stock PyTorch -- the CNN is outside the product -- and replaces the
update lines (``x + a*sign(g)`` / eps-ball projection / [0,1] clamp; 6-14 eager kernels in the
reference) with ONE fused CUDA kernel from libedge_b200.so.  Random numbers are drawn with the same
torch calls in the same order as the reference, so a seeded run follows the same trajectory.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as F_ee


def _random_start(x, epsilon):
    """utils/attacks.py:15-17: x + U(-eps, eps), clamped to [0, 1].  The noise is the reference's own draw
    (zeros_like(x).uniform_ advances the generator identically); add + clamp are one kernel."""
    return F_ee.add_clamp(x, torch.empty_like(x).uniform_(-epsilon, epsilon), 0.0, 1.0)


def _linf_step(x, grad, inputs, step_signed, epsilon):
    """utils/attacks.py:25-27 fused: clamp(min(max(x + a*sign(g), x0-eps), x0+eps), 0, 1)."""
    return F_ee.pgd_linf_step(x.detach(), grad.detach(), inputs.detach(), step_signed, epsilon, 0.0, 1.0)


# Projected Gradient Descent -- utils/attacks.py:12-29
def PGD(model, args, inputs, targets, num_steps, step_size):
    x = inputs.detach()

    if args.random:
        x = _random_start(x, args.epsilon)

    for i in range(num_steps):
        x.requires_grad_()
        with torch.enable_grad():
            logits = model(x)
            loss = F.cross_entropy(logits, targets, reduction='sum')
        grad = torch.autograd.grad(loss, [x])[0]
        x = _linf_step(x, grad, inputs, step_size, args.epsilon)

    return x


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
            logits = self.model(x)
            loss = self.loss_fn(logits, self.y)
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


# targeted PGD with a random target label -- utils/attacks.py:33-56
def targeted_PGD(model, args, inputs, labels, num_steps, step_size, nclass, device):
    x = inputs.detach()
    label_offset = torch.randint(low=1, high=nclass, size=labels.shape).to(device)
    target_labels = torch.fmod(labels + label_offset, nclass)

    if args.random:
        x = _random_start(x, args.epsilon)

    for i in range(num_steps):
        x.requires_grad_()
        with torch.enable_grad():
            logits = model(x)
            loss = F.cross_entropy(logits, target_labels, reduction='sum')
        grad = torch.autograd.grad(loss, [x])[0]
        x = _linf_step(x, grad, inputs, -step_size, args.epsilon)

    return x, target_labels


# utils/attacks.py:59-86
def targeted_PGD_trick(model, args, inputs, labels, num_steps, step_size, nclass, device):
    x = inputs.detach()
    label_offset = torch.randint(low=1, high=nclass, size=labels.shape).to(device)
    target_labels = torch.fmod(labels + label_offset, nclass)

    if args.random:
        init_start = torch.Tensor(x.shape).uniform_(-args.epsilon, args.epsilon).to(device)
        start_from_noise_index = torch.gt(torch.rand([]), args.prob_start_from_clean).type(torch.float32).to(device)
        x = x + start_from_noise_index * init_start
        x = torch.clamp(x, 0.0, 1.0)

    for i in range(num_steps):
        x.requires_grad_()
        with torch.enable_grad():
            logits = model(x)
            loss = F.cross_entropy(logits, target_labels, reduction='sum')
        grad = torch.autograd.grad(loss, [x])[0]
        x = _linf_step(x, grad, inputs, -step_size, args.epsilon)

    return x, target_labels


# utils/attacks.py:89-106
class LabelSmoothLoss(torch.nn.Module):
    def __init__(self, smoothing=0.0):
        super(LabelSmoothLoss, self).__init__()
        self.smoothing = smoothing

    def forward(self, input, target):
        log_prob = F.log_softmax(input, dim=-1)
        weight = input.new_ones(input.size()) * self.smoothing / (input.size(-1) - 1.)
        weight.scatter_(-1, target.unsqueeze(-1), (1. - self.smoothing))
        loss = (-weight * log_prob).sum(dim=-1).mean()
        return loss


def compute_loss_and_error(logits, label, label_smoothing=0.):
    loss_function = LabelSmoothLoss(label_smoothing)
    loss = loss_function(logits, label.long())
    return loss


# FGSM -- utils/attacks.py:110-128 (single signed step + clamp, no eps projection)
def FGSM(model, inputs, target, targeted=False, step_size=0.007):
    x = inputs.detach()
    x.requires_grad_()

    with torch.enable_grad():
        logits = model(x)
        loss = F.cross_entropy(logits, target, reduction='sum')

    grad = torch.autograd.grad(loss, [x])[0]
    step = -step_size if targeted else step_size
    return F_ee.fgsm_step(x.detach(), grad.detach(), step, 0.0, 1.0)


def predict_from_logits(logits, dim=1):
    return logits.max(dim=dim, keepdim=False)[1]


# CW with Linf norm -- utils/attacks.py:136-232
def CWLinfAttack(x, y, model, magnitude, previous_p, max_eps, max_iters=20, target=None, _type='linf',
                 n_class=10, cur_device=None):
    model.eval()
    device = cur_device
    x = x.to(device)
    y = y.to(device)
    if target is not None:
        target = target.to(device)
    adv = x.clone()
    pred = predict_from_logits(model(x))
    if torch.sum((pred == y)).item() == 0:
        return adv, previous_p
    ind_non_suc = (pred == y).nonzero().squeeze()
    x = x[ind_non_suc]
    y = y[ind_non_suc]
    target = target[ind_non_suc]
    x = x if len(x.shape) == 4 else x.unsqueeze(0)
    y = y if len(y.shape) == 1 else y.unsqueeze(0)
    target = target if len(target.shape) == 1 else target.unsqueeze(0)
    if previous_p is not None:
        previous_p = previous_p.to(device)
        previous_p_c = previous_p.clone()
        previous_p = previous_p[ind_non_suc]
        previous_p = previous_p if len(previous_p.shape) == 4 else previous_p.unsqueeze(0)

    one_hot_y = torch.zeros(y.size(0), n_class).to(device)
    one_hot_y[torch.arange(y.size(0)), y] = 1

    # random start
    x.requires_grad = True
    mag = magnitude.item() if isinstance(magnitude, torch.Tensor) else magnitude
    rand_perturb = torch.FloatTensor(x.shape).uniform_(-mag, mag)
    rand_perturb = rand_perturb.to(device)
    adv_imgs = x + rand_perturb
    adv_imgs.clamp_(0, 1)

    if previous_p is not None:
        max_x = x - previous_p + max_eps
        min_x = x - previous_p - max_eps
    else:
        max_x = x + max_eps
        min_x = x - max_eps

    max_iters = int(max_iters)
    x_d, min_d, max_d = x.detach(), min_x.detach(), max_x.detach()

    with torch.enable_grad():
        for _iter in range(max_iters):
            if not adv_imgs.requires_grad:
                adv_imgs.requires_grad_()
            outputs = model(adv_imgs)

            correct_logit = torch.sum(one_hot_y * outputs, dim=1)
            if target is not None:
                wrong_logit = torch.zeros(target.size(0), n_class).to(device)
                wrong_logit[torch.arange(target.size(0)), target] = 1
                wrong_logit = torch.sum(wrong_logit * outputs, dim=1)
            else:
                wrong_logit, _ = torch.max((1 - one_hot_y) * outputs - 1e4 * one_hot_y, dim=1)

            loss = -torch.sum(F.relu(correct_logit - wrong_logit + 50))
            grads = torch.autograd.grad(loss, adv_imgs, grad_outputs=None, only_inputs=True)[0]

            # attacks.py:212-222 fused: step 0.00392, project to x +- magnitude, clamp, project to [min_x, max_x]
            adv_imgs = F_ee.cw_linf_step(adv_imgs.detach(), grads.detach(), x_d, min_d, max_d, 0.00392, mag)

    adv_imgs = adv_imgs.clamp(0, 1)

    now_p = adv_imgs - x_d
    adv[ind_non_suc] = adv_imgs
    if previous_p is not None:
        previous_p_c[ind_non_suc] = previous_p + now_p
        return adv, previous_p_c

    return adv, now_p


def _small_randn_start(x_natural):
    # the reference hard-codes device='cuda' (attacks.py:250,:291,:311,:383,:406); follow the input instead
    return x_natural.detach() + 0.001 * torch.randn(x_natural.shape, device=x_natural.device).detach()


# ALP -- utils/attacks.py:236-272
class ALP:
    def __init__(self, step_size=0.003, epsilon=0.047, perturb_steps=5, beta=1.0):
        self.step_size = step_size
        self.epsilon = epsilon
        self.perturb_steps = perturb_steps
        self.beta = beta

    def reset_steps(self, k):
        self.perturb_steps = k

    def PGD_Linf(self, model, x_natural, y):
        model.eval()
        x_adv = _small_randn_start(x_natural)

        for _ in range(self.perturb_steps):
            x_adv.requires_grad_()
            with torch.enable_grad():
                loss_c = F.cross_entropy(model(x_adv), y)
            grad = torch.autograd.grad(loss_c, [x_adv])[0].detach()
            x_adv = _linf_step(x_adv, grad, x_natural, self.step_size, self.epsilon)

        return x_adv

    def loss(self, model, logits, logits_adv, y, optimizer):
        model.train()
        optimizer.zero_grad()
        loss_robust = 0.5 * F.cross_entropy(logits, y) + 0.5 * F.cross_entropy(logits_adv, y)
        loss_alp = F.mse_loss(logits, logits_adv)
        return loss_robust + self.beta * loss_alp


# Targeted ALP for Tiny ImageNet -- utils/attacks.py:276-333
class targeted_ALP:
    def __init__(self, step_size=0.003, epsilon=0.047, perturb_steps=5, beta=1.0, n_class=200):
        self.step_size = step_size
        self.epsilon = epsilon
        self.perturb_steps = perturb_steps
        self.beta = beta
        self.n_class = n_class

    def reset_steps(self, k):
        self.perturb_steps = k

    def PGD_Linf(self, model, x_natural, y):
        model.eval()
        x_adv = _small_randn_start(x_natural)

        for _ in range(self.perturb_steps):
            x_adv.requires_grad_()
            with torch.enable_grad():
                loss_c = F.cross_entropy(model(x_adv), y)
            grad = torch.autograd.grad(loss_c, [x_adv])[0].detach()
            x_adv = _linf_step(x_adv, grad, x_natural, self.step_size, self.epsilon)

        return x_adv

    def tarPGD_Linf(self, model, x_natural, y, device):
        model.eval()
        label_offset = torch.randint(low=1, high=self.n_class, size=y.shape).to(device)
        target_labels = torch.fmod(y + label_offset, self.n_class)

        x_adv = _small_randn_start(x_natural)

        for _ in range(self.perturb_steps):
            x_adv.requires_grad_()
            with torch.enable_grad():
                loss_c = F.cross_entropy(model(x_adv), target_labels)
            grad = torch.autograd.grad(loss_c, [x_adv])[0].detach()
            x_adv = _linf_step(x_adv, grad, x_natural, -self.step_size, self.epsilon)

        return x_adv

    def loss(self, model, logits, logits_adv, y, optimizer):
        model.train()
        optimizer.zero_grad()
        loss_robust = 0.5 * F.cross_entropy(logits, y) + 0.5 * F.cross_entropy(logits_adv, y)
        loss_alp = F.mse_loss(logits, logits_adv)
        return loss_robust + self.beta * loss_alp


# Targeted ALP for ImageNet -- utils/attacks.py:337-357
def tar_alp_imagenet(model, args, inputs, labels, num_steps, step_size, device):
    x = inputs.detach()
    label_offset = torch.randint(low=1, high=1000, size=labels.shape).to(device)
    target_labels = torch.fmod(labels + label_offset, 1000)

    x = x + 0.001 * torch.randn(x.shape).to(device).detach()

    for i in range(num_steps):
        x.requires_grad_()
        with torch.enable_grad():
            logits = model(x)
            loss = F.cross_entropy(logits, target_labels, reduction='sum')
        grad = torch.autograd.grad(loss, [x])[0]
        x = _linf_step(x, grad, inputs, -step_size, args.epsilon)

    return x, target_labels


# utils/attacks.py:360-366
def squared_l2_norm(x):
    flattened = x.view(x.shape[0], -1)
    return (flattened ** 2).mean(1)


def l2_norm(x):
    return squared_l2_norm(x).sqrt()


# TRADES -- utils/attacks.py:369-429
class Trades:
    def __init__(self, step_size=0.003, epsilon=0.047, perturb_steps=5, beta=1.0):
        self.step_size = step_size
        self.epsilon = epsilon
        self.perturb_steps = perturb_steps
        self.beta = beta
        self.criterion_kl = nn.KLDivLoss(reduction="batchmean")

    def reset_steps(self, k):
        self.perturb_steps = k

    def PGD_L2(self, model, x_natural, logits):
        model.eval()
        x_adv = _small_randn_start(x_natural)
        prob = F.softmax(logits, dim=-1)

        for _ in range(self.perturb_steps):
            with torch.enable_grad():
                x_adv.requires_grad_()
                loss_kl = self.criterion_kl(F.log_softmax(model(x_adv), dim=1), prob)
            grad = torch.autograd.grad(loss_kl, [x_adv])[0].detach()
            # attacks.py:391-399 fused (per-sample RMS norms, step, L2 re-projection, clamp)
            x_adv = F_ee.pgd_l2_step(x_adv.detach(), grad, x_natural.detach(), self.step_size, self.epsilon)

        return x_adv

    def PGD_Linf(self, model, x_natural, logits):
        model.eval()
        x_adv = _small_randn_start(x_natural)
        prob = F.softmax(logits, dim=-1)

        for _ in range(self.perturb_steps):
            x_adv.requires_grad_()
            with torch.enable_grad():
                loss_kl = self.criterion_kl(F.log_softmax(model(x_adv), dim=1), prob)
            grad = torch.autograd.grad(loss_kl, [x_adv])[0].detach()
            x_adv = _linf_step(x_adv, grad, x_natural, self.step_size, self.epsilon)

        return x_adv

    def loss(self, model, logits, x_adv, labels, optimizer):
        model.train()
        optimizer.zero_grad()
        prob = F.softmax(logits, dim=-1)
        loss_natural = F.cross_entropy(logits, labels)
        loss_robust = self.criterion_kl(F.log_softmax(model(x_adv), dim=1), prob)
        return loss_natural + self.beta * loss_robust


# AVmixup -- utils/attacks.py:433-518
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
        x = inputs.detach()
        if self.args.random:
            x = _random_start(x, self.args.epsilon)
        for i in range(self.num_steps):
            x.requires_grad_()
            with torch.enable_grad():
                logits = model(x)
                log_prob = F.log_softmax(logits, dim=1)
                loss = -torch.sum(log_prob * soft_targets)
            grad = torch.autograd.grad(loss, [x])[0]
            x = _linf_step(x, grad, inputs, step_signed, self.args.epsilon)
        return x

    def _mix(self, x, inputs, targets):
        y_nat = self._label_smoothing(targets, self.lambda1)
        y_vertex = self._label_smoothing(targets, self.lambda2)
        x_weight = np.random.beta(1.0, 1.0, [x.shape[0], 1, 1, 1])
        x_weight_torch = torch.from_numpy(x_weight).to(x.device)
        y_weight = torch.from_numpy(np.reshape(x_weight, [-1, 1])).to(x.device)
        # attacks.py:469-471 + :476 (vertex, clamp, float64 mix, cast back) fused into one pass
        x = F_ee.avmixup_mix(x.detach(), inputs.detach(), x_weight_torch, self.gamma)
        y = y_nat * y_weight + y_vertex * (1 - y_weight)
        return x, y

    def perturb(self, model, inputs, targets):
        """attacks.py:447-479 (targets are soft / one-hot labels)."""
        x = self._attack(model, inputs, targets, self.step_size)
        return self._mix(x, inputs, targets)

    def tar_perturb(self, model, inputs, targets):
        """attacks.py:481-518: descends towards random target labels."""
        label_offset = torch.randint(low=1, high=self.num_classes, size=targets.shape).to(self.device)
        target_labels = torch.fmod(targets + label_offset, self.num_classes)
        x = self._attack(model, inputs, target_labels, -self.step_size)
        return self._mix(x, inputs, targets)


# free / fast adversarial training noise update (in-script loops of the reference):
#   ImageNet/free_imagenet/AT_hfs_canny_free_imagenet_ddp.py:312-315,:330-332
#   ImageNet/fgsm_imagenet/main_fast.py:233-235,:246-253 ; lib/utils.py:36-37
def free_at_update_(global_noise, noise_grad, inputs, fgsm_step, clip_eps):
    """global_noise[0:B] += fgsm_step*sign(noise_grad); clamp to +-clip_eps (in place) and return the
    next repeat's input clamp(inputs + global_noise[0:B], 0, 1), all in one kernel."""
    B = inputs.size(0)
    view = global_noise[0:B]
    return F_ee.free_at_step_(view, noise_grad.detach(), inputs.detach(), fgsm_step, clip_eps, 0.0, 1.0, True)
