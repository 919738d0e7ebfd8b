"""Launch the native HighFreqSuppress kernel a few times (for ncu).  usage: python tools/prof_hfs.py [B] [N] [r]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edge_enhancement_b200 import functional as F  # noqa: E402
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = int(sys.argv[2]) if len(sys.argv) > 2 else 64
r = int(sys.argv[3]) if len(sys.argv) > 3 else 8
x = torch.rand(B, 3, N, N, device="cuda"); y = torch.empty_like(x)
for _ in range(3):
    F.hfs(x, r, out=y)
torch.cuda.synchronize()
print("ok", float(y.sum()))
