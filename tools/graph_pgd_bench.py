"""attacks.PGD (eager launches) vs attacks.GraphedPGD (one captured iteration replayed) at the reference's batch sizes,
where the PGD loop is launch-bound.  BASELINE configs[0]: MNIST small CNN, edge-enhanced PGD-40, batch 128, 1x28x28, full
CannyFilter (alpha 0.3, low/high 25/51); configs[1]-like: 256x3x64x64, step125, PGD-10, a small conv net.
usage: python tools/graph_pgd_bench.py"""
import contextlib
import io
import json
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edge_enhancement_b200 import attacks, core  # noqa: E402

dev = "cuda:0"


class SmallNet(nn.Module):
    """two conv + two linear layers, the size of the reference's MNIST Net2"""

    def __init__(self, c, side, canny, low, high, n_class=10):
        super().__init__()
        self.canny, self.low, self.high = canny, low, high
        self.conv1, self.conv2 = nn.Conv2d(c, 32, 5, padding=2), nn.Conv2d(32, 64, 5, padding=2)
        self.fc1, self.fc2 = nn.Linear(64 * (side // 4) ** 2, 256), nn.Linear(256, n_class)

    def forward(self, x):
        x = core.edge_enhance(x, x, self.canny, 1.0, self.low, self.high, True)
        x = F.max_pool2d(F.relu(self.conv1(x)), 2)
        x = F.max_pool2d(F.relu(self.conv2(x)), 2)
        return self.fc2(F.relu(self.fc1(x.flatten(1))))


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for name, shape, filt, alpha, low, high, eps, step, steps in (
        ("configs[0] MNIST 128x1x28x28, CannyFilter, PGD-40", (128, 1, 28, 28), core.CannyFilter, 0.3, 25 / 255, 51 / 255, 0.3, 0.01, 40),
        ("configs[1]-like 256x3x64x64, step125, PGD-10", (256, 3, 64, 64), core.CannyFilter_step125_1, 0.0, None, 76 / 255, 16 / 255, 2 / 255, 10)):
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        canny = filt(use_cuda=False, alpha=alpha)
    model = SmallNet(shape[1], shape[2], canny, low, high).to(dev).eval()
    x = torch.rand(shape, device=dev)
    y = torch.randint(0, 10, (shape[0],), device=dev)

    class A:
        random = False
        epsilon = eps

    eager = lambda: attacks.PGD(model, A, x, y, steps, step)
    g = attacks.GraphedPGD(model, A, x, y, step)
    graphed = lambda: g(x, y, steps)
    same = bool(torch.equal(eager(), graphed()))
    te, tg = timeit(eager), timeit(graphed)
    print(json.dumps({"config": name, "eager_ms_per_attack": te, "graphed_ms_per_attack": tg, "speedup": te / tg,
                      "images_per_s_eager": shape[0] / te * 1e3, "images_per_s_graphed": shape[0] / tg * 1e3,
                      "bit_identical": same}), flush=True)
