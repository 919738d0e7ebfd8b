"""Edge-enhanced PGD-10 adversarial TRAINING throughput (images/s) with a real CNN, 1..N GPUs (DDP over NCCL).

    python tools/train_throughput.py [--front ours|eager|both] [--iters 6] [--batch 256]
    torchrun --nproc-per-node N tools/train_throughput.py ...

BASELINE.json configs[1]: Tiny-ImageNet PreAct-ResNet18, 3x64x64, batch 256 per GPU (weak scaling), PGD-10
(eps 16/255, step 2/255), CannyFilter_step125_1 front end (high 76/255, w = 1, r = 8), SGD.  The CNN, the loss, DDP's
gradient all-reduce and the FFT low-pass stay on stock PyTorch (outside the product); what changes between the two
front ends is ONLY the hot path:
  ours  : core.EdgeEnhance (fused edge + blend kernels) and attacks.PGD (fused sign/project/clamp step)
  eager : the same mathematics composed from stock torch ops the way the reference composes it (per-channel
          replicate-pad + 3x3 conv, channel-summing Sobel convs, pow / where / masked writes, autograd through all
          of it; 8 elementwise kernels per PGD update; the five host-side zeros + H2D scratch allocations per call) --
          written here from SURVEY.md section 2.3, it stands for "the reference's op chain on the same B200" (the
          reference itself cannot travel to the GPU box).
  eager_clean : the same without the reference's dead host-side allocations (what a tidy eager version would cost).
One iteration = PGD-10 attack (10 x model forward + input gradient + update) + training forward/backward + SGD step,
i.e. what Tiny_ImageNet/experiments_tinyimagenet.py:train does per batch.  Synthetic data, random-init weights.
Prints one JSON line per front end (rank 0).  This is supporting evidence for the north-star metric; the judged
bench line is bench.py (hot path only).
"""
import argparse
import contextlib
import io
import json
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

EPS, ALPHA, HIGH = 16 / 255, 2 / 255, 76 / 255


# ---- stock CNN: PreAct-ResNet18 for 64x64 inputs (AWP/Tiny_imagenet/models_tiny_awp/preactresnet.py is the reference's) ----
class PreActBlock(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.bn1, self.bn2 = nn.BatchNorm2d(cin), nn.BatchNorm2d(cout)
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.short = nn.Conv2d(cin, cout, 1, stride, bias=False) if (stride != 1 or cin != cout) else None

    def forward(self, x):
        o = F.relu(self.bn1(x))
        sc = self.short(o) if self.short is not None else x
        o = self.conv1(o)
        o = self.conv2(F.relu(self.bn2(o)))
        return o + sc


class PreActResNet18(nn.Module):
    def __init__(self, front, n_class=200):
        super().__init__()
        self.front = front
        self.conv1 = nn.Conv2d(3, 64, 3, 1, 1, bias=False)
        cfg, layers, cin = [(64, 1), (128, 2), (256, 2), (512, 2)], [], 64
        for cout, stride in cfg:
            layers += [PreActBlock(cin, cout, stride), PreActBlock(cout, cout, 1)]
            cin = cout
        self.layers = nn.Sequential(*layers)
        self.bn = nn.BatchNorm2d(512)
        self.fc = nn.Linear(512, n_class)

    def forward(self, x):
        x = self.front(x)
        o = self.layers(self.conv1(x))
        o = F.relu(self.bn(o))
        return self.fc(F.adaptive_avg_pool2d(o, 1).flatten(1))


class Bottleneck(nn.Module):
    def __init__(self, cin, mid, stride):
        super().__init__()
        cout = mid * 4
        self.c1, self.b1 = nn.Conv2d(cin, mid, 1, bias=False), nn.BatchNorm2d(mid)
        self.c2, self.b2 = nn.Conv2d(mid, mid, 3, stride, 1, bias=False), nn.BatchNorm2d(mid)
        self.c3, self.b3 = nn.Conv2d(mid, cout, 1, bias=False), nn.BatchNorm2d(cout)
        self.down = None
        if stride != 1 or cin != cout:
            self.down = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        o = F.relu(self.b1(self.c1(x)))
        o = F.relu(self.b2(self.c2(o)))
        o = self.b3(self.c3(o))
        return F.relu(o + (x if self.down is None else self.down(x)))


class ResNet50(nn.Module):
    """stock ResNet-50 (ImageNet/models_imagenet/resnet_EE.py is the reference's), front end applied to the input"""

    def __init__(self, front, n_class=1000):
        super().__init__()
        self.front = front
        self.conv1, self.bn1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False), nn.BatchNorm2d(64)
        layers, cin = [], 64
        for mid, n, stride in ((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)):
            for i in range(n):
                layers.append(Bottleneck(cin, mid, stride if i == 0 else 1))
                cin = mid * 4
        self.layers = nn.Sequential(*layers)
        self.fc = nn.Linear(2048, n_class)

    def forward(self, x):
        x = self.front(x)
        o = F.max_pool2d(F.relu(self.bn1(self.conv1(x))), 3, 2, 1)
        o = self.layers(o)
        return self.fc(F.adaptive_avg_pool2d(o, 1).flatten(1))


# ---- eager front end: the reference's composition out of stock torch ops ----
class _ThresholdSTE(torch.autograd.Function):
    """binary mask `v > thr` with the straight-through window (thr, 1.001] in the backward"""

    @staticmethod
    def forward(ctx, v, thr):
        ctx.save_for_backward(v, thr)
        out = v.clone()
        out[v > thr] = 1.0
        out[v <= thr] = 0.0
        return out

    @staticmethod
    def backward(ctx, g):
        v, thr = ctx.saved_tensors
        gi = g.clone()
        gi[v > 1.001] = 0
        gi[v <= thr] = 0
        return gi, None


class EagerFront(nn.Module):
    def __init__(self, size, r, w, high, dev, faithful=True):
        super().__init__()
        self.faithful = faithful        # True: also the reference's per-call host-side zeros + H2D scratch allocations
        import numpy as np
        from edge_enhancement_b200 import core
        self.w, self.high = w, high
        self.hfs = core.HighFreqSuppress(size, size, r, impl='torch_fft')   # the reference-style arm keeps the torch.fft low-pass
        g = torch.from_numpy(core.get_gaussian_kernel(3, 0, 1)).float()[None, None].to(dev)
        sx = torch.from_numpy(core.get_sobel_kernel(3)).float()[None, None].to(dev)
        self.wg, self.wsx, self.wsy = g, sx, sx.transpose(2, 3).contiguous()
        self.pad = nn.ReplicationPad2d(1)

    def edge(self, img):
        B, C, H, W = img.shape
        dev = img.device
        if self.faithful:               # utils/core.py:553-557: five torch.zeros(...).to(device) per call, four of them unused
            blurred = torch.zeros((B, C, H, W)).to(dev)
            for _name in ("gx", "gy", "mag", "ori"):
                torch.zeros((B, 1, H, W)).to(dev)
        else:
            blurred = torch.empty((B, C, H, W), device=dev)
        for c in range(C):
            blurred[:, c:c + 1] = F.conv2d(self.pad(img[:, c:c + 1]), self.wg)
        pb = self.pad(blurred)
        gx = F.conv2d(pb, self.wsx.repeat(1, C, 1, 1)) / C
        gy = F.conv2d(pb, self.wsy.repeat(1, C, 1, 1)) / C
        mag = (gx ** 2 + gy ** 2) ** 0.5
        mag = torch.where(mag < 0.0, torch.zeros_like(mag), mag)          # alpha gate (alpha = 0)
        thin = mag.clone()
        return _ThresholdSTE.apply(thin, torch.tensor(self.high)) * 1

    def forward(self, x):
        x_hfs = self.hfs(x)
        x = x_hfs + self.w * self.edge(x)
        return torch.clamp(x, 0.0, 1.0)


def eager_pgd(model, x0, y, steps, alpha, eps):
    x = x0.detach()
    for _ in range(steps):
        x.requires_grad_()
        with torch.enable_grad():
            loss = F.cross_entropy(model(x), y, reduction='sum')
        grad = torch.autograd.grad(loss, [x])[0]
        x = x.detach() + alpha * torch.sign(grad.detach())
        x = torch.min(torch.max(x, x0 - eps), x0 + eps)
        x = torch.clamp(x, 0, 1)
    return x


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--front", default="all", choices=["ours", "eager", "eager_clean", "all"])
    ap.add_argument("--iters", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--batch", type=int, default=256, help="images per GPU")
    ap.add_argument("--side", type=int, default=64)
    ap.add_argument("--hfs-impl", default="native", choices=["native", "tcgen05"],
                    help="low-pass kernel of the `ours` front end (core.set_hfs_impl): FFMA (exact) or tensor cores (64 px only)")
    ap.add_argument("--pgd-steps", type=int, default=10)
    ap.add_argument("--config", default="tiny_pgd", choices=["tiny_pgd", "imagenet_free"],
                    help="tiny_pgd: BASELINE configs[1] (default); imagenet_free: configs[3], ResNet-50, 3x224x224, 32 images per GPU, "
                         "free adversarial training with n_repeats = 4, clip_eps = fgsm_step = 4/255 "
                         "(ImageNet/free_imagenet/AT_hfs_canny_free_imagenet_ddp.py:288-334)")
    args = ap.parse_args()
    free = (args.config == "imagenet_free")
    if free:
        args.side, args.batch = 224, (32 if args.batch == 256 else args.batch)
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from edge_enhancement_b200 import attacks, core
    core.set_hfs_impl(args.hfs_impl)

    class A:
        random = True
        epsilon = EPS

    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = torch.rand((args.batch, 3, args.side, args.side), device=dev, generator=gen)
    y = torch.randint(0, 200, (args.batch,), device=dev, generator=gen)
    results = []
    for front_name in (["ours", "eager", "eager_clean"] if args.front == "all" else [args.front]):
        torch.manual_seed(0)
        if front_name == "ours":
            with contextlib.redirect_stdout(io.StringIO()):
                front = core.EdgeEnhance(cize=args.side, r=16 if free else 8, w=1.0, low=38.0, high=76.0, alpha=0.0, sigma=1,
                                         type_canny='CannyFilter_step125_1')
        else:
            front = EagerFront(args.side, 16 if free else 8, 1.0, HIGH, dev, faithful=(front_name == "eager"))
        model = (ResNet50(front) if free else PreActResNet18(front)).to(dev)
        if free and world > 1:
            model = nn.SyncBatchNorm.convert_sync_batchnorm(model)          # as the reference does (:170)
        net = nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
        opt = torch.optim.SGD(net.parameters(), lr=0.1, momentum=0.9, weight_decay=2e-4)

        noise = torch.zeros((args.batch, 3, args.side, args.side), device=dev)      # the reference's global_noise_data
        clip_eps = fgsm_step = 4 / 255

        def free_iteration():
            """one mini-batch of free adversarial training: n_repeats = 4 replays, each a forward/backward that updates the
            weights AND the persistent noise (AT_hfs_canny_free_imagenet_ddp.py:311-334)"""
            net.train()
            in1 = torch.clamp(x + noise, 0, 1.0)
            for _ in range(4):
                in1 = in1.detach().requires_grad_()
                loss = F.cross_entropy(net(in1), y)
                opt.zero_grad(set_to_none=True)
                loss.backward()
                if front_name == "ours":
                    in1 = attacks.free_at_update_(noise, in1.grad, x, fgsm_step, clip_eps)      # delta update + next input, one kernel
                else:
                    noise.add_(fgsm_step * torch.sign(in1.grad))
                    noise.clamp_(-clip_eps, clip_eps)
                    in1 = x + noise
                    in1.clamp_(0, 1.0)
                opt.step()
            return loss

        def iteration():
            if free:
                return free_iteration()
            net.eval()                                       # attack in eval mode keeps BN statistics out of the inner loop
            if front_name == "ours":
                x_adv = attacks.PGD(net, A, x, y, args.pgd_steps, ALPHA)
            else:
                xs = torch.clamp(x + torch.zeros_like(x).uniform_(-EPS, EPS), 0, 1)
                x_adv = eager_pgd(net, xs, y, args.pgd_steps, ALPHA, EPS)
            net.train()
            loss = F.cross_entropy(net(x_adv), y)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            return loss

        for _ in range(args.warmup):
            iteration()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            loss = iteration()
        e1.record()
        torch.cuda.synchronize(dev)
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ips = world * args.batch * args.iters / (float(ms.item()) / 1e3)
        results.append((front_name, ips))
        if rank == 0:
            print(json.dumps({"metric": ("edge-enhanced free adversarial training (n_repeats 4) images/sec" if free else
                                         "edge-enhanced PGD-%d adversarial training images/sec" % args.pgd_steps), "front_end": front_name,
                              "value": ips, "unit": "images/s", "n_gpus": world, "ms_per_iteration": float(ms.item()) / args.iters,
                              "config": "%s (stock torch ops, fp32, cudnn TF32 default), 3x%dx%d, batch %d per GPU, "
                                        "CannyFilter_step125_1 + %s low-pass, %s, SGD; DDP/NCCL for N > 1"
                                        % ("ResNet-50 + SyncBN" if free else "PreAct-ResNet18", args.side, args.side, args.batch,
                                           ("native (%s)" % args.hfs_impl) if front_name == "ours" else "torch.fft",
                                           "clip_eps = fgsm_step = 4/255, 4 replays per batch" if free else "eps 16/255, step 2/255"),
                              "final_loss": float(loss.item()), "data": "synthetic"}), flush=True)
        del net, model, opt
        torch.cuda.empty_cache()
    if rank == 0 and len(results) > 1:
        print(json.dumps({"speedup_ours_over_" + n: results[0][1] / v for n, v in results[1:]}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
