import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from edge_enhancement_b200 import functional as F
from oracle import oracle as O
for N, r in ((64, 8), (28, 4), (32, 8), (128, 12)):
    x = np.random.default_rng(N).standard_normal((2, 3, N, N)).astype(np.float32)
    got = F.hfs(torch.from_numpy(x).cuda(), r).cpu().numpy()
    want = O.hfs(x, r)
    bad = got != want
    print(N, r, "mismatch", bad.sum(), "of", bad.size, "max diff", np.abs(got - want).max())
    if bad.any():
        idx = np.argwhere(bad)
        cols = np.bincount(idx[:, 3], minlength=N)
        rows = np.bincount(idx[:, 2], minlength=N)
        print("  by column:", cols.tolist())
        print("  by row   :", rows.tolist())
