"""Shared helpers for the parity tests: seeded synthetic inputs and the
CUDA-side call wrappers (everything goes through the C ABI via edge_enhancement_b200.functional)."""
import numpy as np

HIGH = 76.0 / 255.0
LOW = 38.0 / 255.0


def rng(seed):
    return np.random.default_rng(seed)


def make_inputs(seed, B, C, H, W, kind="uniform"):
    """x, base, g_out, g_edge as float32 numpy arrays."""
    r = rng(seed)
    if kind == "uniform":
        x = r.random((B, C, H, W), dtype=np.float32)
    elif kind == "sparse":           # MNIST-like: ~80 % exact zeros (flat regions, mag == 0)
        x = np.maximum(r.random((B, C, H, W), dtype=np.float32) - 0.8, 0).astype(np.float32) * 5
    elif kind == "smooth":           # low-frequency content: long edges, fewer isolated maxima
        yy, xx = np.meshgrid(np.linspace(0, 3, H), np.linspace(0, 3, W), indexing="ij")
        ph = r.random((B, C, 1, 1)) * 6.28
        x = (0.5 + 0.5 * np.sin(2.1 * yy[None, None] + ph) * np.cos(1.7 * xx[None, None] - ph)).astype(np.float32)
    else:
        raise ValueError(kind)
    base = (r.random((B, C, H, W), dtype=np.float32) * 1.1 - 0.1).astype(np.float32)
    g_out = r.standard_normal((B, C, H, W), dtype=np.float32)
    g_edge = r.standard_normal((B, 1, H, W), dtype=np.float32)
    return x, base, g_out, g_edge


def make_attack_inputs(seed, shape, eps):
    r = rng(seed)
    x0 = r.random(shape, dtype=np.float32)
    x = np.clip(x0 + (r.random(shape, dtype=np.float32) * 2 - 1).astype(np.float32) * np.float32(eps), 0, 1).astype(np.float32)
    g = r.standard_normal(shape, dtype=np.float32)
    flat = g.reshape(-1)
    idx = r.choice(flat.size, size=max(1, flat.size // 100), replace=False)
    flat[idx] = 0.0                                    # pins sign(0) = 0
    if flat.size > 8:
        flat[1] = np.float32(-0.0)
        flat[2] = np.float32(1e-30)
        flat[3] = np.float32(-1e-30)
    return x, g, x0


MODES = {
    # name: (low, high, hysteresis)
    "hyst": (LOW, HIGH, True),
    "mix": (LOW, HIGH, False),
    "low": (LOW, None, False),
    "raw": (None, None, False),
}
