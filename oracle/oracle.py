"""numpy/ctypes front-end of the C oracle (oracle/ee_oracle.c).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  All functions take and return
C-contiguous float32 numpy arrays in NCHW; nothing here touches CUDA or torch.
"""
import ctypes
import os

import numpy as np

from . import build as _build

STEP125, CANNY, BPDA = 0, 1, 2
VARIANTS = {"step125": STEP125, "canny": CANNY, "bpda": BPDA}


class OracleParams(ctypes.Structure):
    _fields_ = [("variant", ctypes.c_int),
                ("c0", ctypes.c_float), ("c1", ctypes.c_float), ("c2", ctypes.c_float),
                ("alpha", ctypes.c_float),
                ("low_thr", ctypes.c_float), ("high_thr", ctypes.c_float),
                ("has_low", ctypes.c_int), ("has_high", ctypes.c_int),
                ("hysteresis", ctypes.c_int), ("nan_compat", ctypes.c_int)]


def gaussian3(mu=0.0, sigma=1.0):
    """fp32 3x3 Gaussian exactly as utils/core.py:58-72 + the .type(torch.float) of :164."""
    g1 = np.linspace(-1, 1, 3)
    x, y = np.meshgrid(g1, g1)
    d = (x ** 2 + y ** 2) ** 0.5
    g = np.exp(-(d - mu) ** 2 / (2 * sigma ** 2)) / (2 * np.pi * sigma ** 2)
    g = g / np.sum(g)
    return g.astype(np.float32)


def make_params(variant="step125", mu=0.0, sigma=1.0, alpha=0.0, low=None, high=None, hysteresis=False, nan_compat=False):
    g = gaussian3(mu, sigma)
    v = VARIANTS[variant] if isinstance(variant, str) else int(variant)
    return OracleParams(v, float(g[0, 0]), float(g[0, 1]), float(g[1, 1]), float(np.float32(alpha)),
                        float(np.float32(0.0 if low is None else low)),
                        float(np.float32(0.0 if high is None else high)),
                        int(low is not None), int(high is not None), int(bool(hysteresis)), int(bool(nan_compat)))


_LIB = None


def _cpu_has_fma():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    fl = line.split()
                    return "fma" in fl and "avx2" in fl
    except OSError:
        pass
    return False


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    kind = "fma" if _cpu_has_fma() else "generic"
    path = _build.lib_path(kind)
    if not os.path.exists(path):
        _build.build()
    L = ctypes.CDLL(path)
    fp = ctypes.POINTER(ctypes.c_float)
    pp = ctypes.POINTER(OracleParams)
    i, i64, f = ctypes.c_int, ctypes.c_int64, ctypes.c_float
    L.ee_oracle_edge_fwd.argtypes = [fp, fp, i, i, i, i, pp]
    L.ee_oracle_edge_fwd.restype = i
    L.ee_oracle_edge_blend_fwd.argtypes = [fp, fp, fp, fp, i, i, i, i, pp, f]
    L.ee_oracle_edge_blend_fwd.restype = i
    L.ee_oracle_edge_bwd.argtypes = [fp, fp, fp, i, i, i, i, pp]
    L.ee_oracle_edge_bwd.restype = i
    L.ee_oracle_edge_blend_bwd.argtypes = [fp, fp, fp, fp, fp, i, i, i, i, pp, f]
    L.ee_oracle_edge_blend_bwd.restype = i
    L.ee_oracle_pgd_linf_step.argtypes = [fp, fp, fp, fp, i64, f, f, f, f]
    L.ee_oracle_pgd_linf_step.restype = None
    L.ee_oracle_fgsm_step.argtypes = [fp, fp, fp, i64, f, f, f]
    L.ee_oracle_fgsm_step.restype = None
    L.ee_oracle_free_at_step.argtypes = [fp, fp, fp, fp, i64, f, f, f, f]
    L.ee_oracle_free_at_step.restype = None
    L.ee_oracle_cw_linf_step.argtypes = [fp, fp, fp, fp, fp, fp, i64, f, f]
    L.ee_oracle_cw_linf_step.restype = None
    L.ee_oracle_pgd_l2_step.argtypes = [fp, fp, fp, fp, i, i64, f, f, i]
    L.ee_oracle_pgd_l2_step.restype = None
    L.ee_oracle_gf_blend_fwd.argtypes = [fp, fp, fp, i, i, i, i, f, f, f, f]
    L.ee_oracle_gf_blend_fwd.restype = None
    L.ee_oracle_gf_blend_bwd.argtypes = [fp, fp, fp, fp, fp, i, i, i, i, f, f, f, f]
    L.ee_oracle_gf_blend_bwd.restype = i
    L.ee_oracle_add_clamp.argtypes = [fp, fp, fp, i64, f, f]
    L.ee_oracle_add_clamp.restype = None
    L.ee_oracle_avmixup_mix.argtypes = [fp, fp, ctypes.POINTER(ctypes.c_double), fp, i, i64, f]
    L.ee_oracle_avmixup_mix.restype = None
    L.ee_oracle_hfs.argtypes = [fp, fp, fp, i, i, i, fp, fp, fp, f, i, i, i]
    L.ee_oracle_hfs.restype = i
    L.ee_oracle_add_square.argtypes = [fp, fp, fp, fp, fp, i, i, i, i, i, f]
    L.ee_oracle_add_square.restype = None
    for name in ("to_compare",):
        getattr(L, "ee_oracle_%s_fwd" % name).argtypes = [fp, fp, i64, f]
        getattr(L, "ee_oracle_%s_bwd" % name).argtypes = [fp, fp, fp, i64, f]
    for name in ("to_eq", "safe_sign"):
        getattr(L, "ee_oracle_%s_fwd" % name).argtypes = [fp, fp, i64]
        getattr(L, "ee_oracle_%s_bwd" % name).argtypes = [fp, fp, fp, i64]
    _LIB = L
    return L


def threads():
    """Host threads the OpenMP loops of the oracle use."""
    L = lib()
    L.ee_oracle_threads.restype = ctypes.c_int
    return int(L.ee_oracle_threads())


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _chk(rc):
    if rc != 0:
        raise MemoryError("oracle allocation failed")


def edge_fwd(x, params):
    x = _f32(x)
    B, C, H, W = x.shape
    edge = np.empty((B, 1, H, W), np.float32)
    _chk(lib().ee_oracle_edge_fwd(_p(x), _p(edge), B, C, H, W, ctypes.byref(params)))
    return edge


def edge_blend_fwd(x, base, params, w, want_edge=False):
    x, base = _f32(x), _f32(base)
    B, C, H, W = x.shape
    out = np.empty_like(x)
    edge = np.empty((B, 1, H, W), np.float32) if want_edge else None
    _chk(lib().ee_oracle_edge_blend_fwd(_p(x), _p(base), _p(out), _p(edge), B, C, H, W,
                                        ctypes.byref(params), float(np.float32(w))))
    return (out, edge) if want_edge else out


def edge_bwd(g_edge, x, params):
    g_edge, x = _f32(g_edge), _f32(x)
    B, C, H, W = x.shape
    g_x = np.empty_like(x)
    _chk(lib().ee_oracle_edge_bwd(_p(g_edge), _p(x), _p(g_x), B, C, H, W, ctypes.byref(params)))
    return g_x


def edge_blend_bwd(g_out, x, base, params, w):
    g_out, x, base = _f32(g_out), _f32(x), _f32(base)
    B, C, H, W = x.shape
    g_x, g_base = np.empty_like(x), np.empty_like(x)
    _chk(lib().ee_oracle_edge_blend_bwd(_p(g_out), _p(x), _p(base), _p(g_x), _p(g_base), B, C, H, W,
                                        ctypes.byref(params), float(np.float32(w))))
    return g_x, g_base


def pgd_linf_step(x, g, x0, alpha_signed, eps, lo=0.0, hi=1.0):
    x, g, x0 = _f32(x), _f32(g), _f32(x0)
    out = np.empty_like(x)
    lib().ee_oracle_pgd_linf_step(_p(x), _p(g), _p(x0), _p(out), x.size, alpha_signed, eps, lo, hi)
    return out


def fgsm_step(x, g, alpha_signed, lo=0.0, hi=1.0):
    x, g = _f32(x), _f32(g)
    out = np.empty_like(x)
    lib().ee_oracle_fgsm_step(_p(x), _p(g), _p(out), x.size, alpha_signed, lo, hi)
    return out


def free_at_step(delta, g, x0, alpha, eps, lo=0.0, hi=1.0):
    """Returns (delta_new, x_adv); does not modify its inputs."""
    d = _f32(delta).copy()
    g, x0 = _f32(g), _f32(x0)
    x_adv = np.empty_like(x0)
    lib().ee_oracle_free_at_step(_p(d), _p(g), _p(x0), _p(x_adv), d.size, alpha, eps, lo, hi)
    return d, x_adv


def cw_linf_step(adv, g, x, min_x, max_x, step, magnitude):
    adv, g, x, min_x, max_x = map(_f32, (adv, g, x, min_x, max_x))
    out = np.empty_like(adv)
    lib().ee_oracle_cw_linf_step(_p(adv), _p(g), _p(x), _p(min_x), _p(max_x), _p(out), adv.size, step, magnitude)
    return out


def pgd_l2_step(x, g, x0, step, eps, three_pass=False):
    """three_pass=True: the reduction order of the any-shape three-pass kernel instead of the cluster kernel's."""
    x, g, x0 = _f32(x), _f32(g), _f32(x0)
    out = np.empty_like(x)
    B = x.shape[0]
    lib().ee_oracle_pgd_l2_step(_p(x), _p(g), _p(x0), _p(out), B, x.size // B, step, eps, 0 if three_pass else -1)
    return out


def gf_blend_fwd(edge, base, w, gauss=None):
    """clamp(base + w * conv2d(edge, gauss, padding=1), 0, 1) -- the with_gf=True blend (resnet_EE.py:185-191)."""
    edge, base = _f32(edge), _f32(base)
    g = gaussian3() if gauss is None else np.asarray(gauss, np.float32)
    B, C, H, W = base.shape
    out = np.empty_like(base)
    lib().ee_oracle_gf_blend_fwd(_p(edge), _p(base), _p(out), B, C, H, W, float(g[0, 0]), float(g[0, 1]), float(g[1, 1]), w)
    return out


def gf_blend_bwd(g_out, edge, base, w, gauss=None):
    g_out, edge, base = _f32(g_out), _f32(edge), _f32(base)
    g = gaussian3() if gauss is None else np.asarray(gauss, np.float32)
    B, C, H, W = base.shape
    g_edge, g_base = np.empty_like(edge), np.empty_like(base)
    _chk(lib().ee_oracle_gf_blend_bwd(_p(g_out), _p(edge), _p(base), _p(g_edge), _p(g_base), B, C, H, W,
                                      float(g[0, 0]), float(g[0, 1]), float(g[1, 1]), w))
    return g_edge, g_base


def add_clamp(x, noise, lo=0.0, hi=1.0):
    x, noise = _f32(x), _f32(noise)
    out = np.empty_like(x)
    lib().ee_oracle_add_clamp(_p(x), _p(noise), _p(out), x.size, lo, hi)
    return out


def avmixup_mix(x_adv, inputs, weight, gamma):
    x_adv, inputs = _f32(x_adv), _f32(inputs)
    w = np.ascontiguousarray(weight, dtype=np.float64).reshape(-1)
    B = x_adv.shape[0]
    out = np.empty_like(x_adv)
    lib().ee_oracle_avmixup_mix(_p(x_adv), _p(inputs), w.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), _p(out), B,
                                x_adv.size // B, float(np.float32(gamma)))
    return out


def hfs_tables(N, r):
    """Real Fourier bases and mixing weights of HighFreqSuppress(N, N, r) (float64 -> float32), shared by the oracle, the
    CUDA kernel and core.HighFreqSuppress: cb[N][NJp] = 1, cos(k th), sin(k th) (k < r); rb[N][NIp] = 1, cos, sin (k <= r);
    w[NIp][NJp] = alpha_i * beta_j with alpha = (1, 2.., 1 at k = r)/N, beta = (1, 2..)/N; gamma = 2/N^2."""
    n = np.arange(N)
    th = 2 * np.pi * n / N
    NJ, NI = 2 * r - 1, 2 * r + 1
    NJp, NIp = (NJ + 3) // 4 * 4, (NI + 3) // 4 * 4
    cb = np.zeros((N, NJp)); rb = np.zeros((N, NIp))
    cb[:, 0] = 1.0; rb[:, 0] = 1.0
    for k in range(1, r):
        cb[:, k] = np.cos(k * th); cb[:, r - 1 + k] = np.sin(k * th)
    for k in range(1, r + 1):
        rb[:, k] = np.cos(k * th); rb[:, r + k] = np.sin(k * th)
    if N % 2 == 0:
        cb[N // 2, r:] = 0.0           # sin(k pi) is exactly 0 (the folded kernel's sine chains start at w = N/2)
    alpha = np.zeros(NIp); beta = np.zeros(NJp)
    alpha[0] = 1.0 / N; alpha[1:r] = 2.0 / N; alpha[r] = 1.0 / N; alpha[r + 1:2 * r] = 2.0 / N; alpha[2 * r] = 1.0 / N
    beta[0] = 1.0 / N; beta[1:NJ] = 2.0 / N
    w = np.outer(alpha, beta)
    return (np.ascontiguousarray(cb, np.float32), np.ascontiguousarray(rb, np.float32),
            np.ascontiguousarray(w, np.float32), float(np.float32(2.0 / (N * N))))


def hfs(x, r, add=None):
    """HighFreqSuppress(N, N, r) on [..., N, N] planes (C restatement of the five-product form); + add if given."""
    x = _f32(x)
    add = None if add is None else _f32(add)
    N = x.shape[-1]
    assert x.shape[-2] == N
    cb, rb, w, gamma = hfs_tables(N, r)
    y = np.empty_like(x)
    # summation trees of the kernel that serves this shape (ee_hfs.cuh): whole-plane kernel 1 / 1; row-blocked kernel for
    # large planes: stage 1 over 8 lanes (4 when 2r-1 > 32), stage 2 over 2 lanes
    # 224 px: half-plane blocks, no lane split in stage 1; 288 px: 8-row blocks with the lane split
    rbk = 8 if N >= 256 else 16
    ks1, ks2 = (1, 1) if N <= 128 else ((1 if N == 224 else (8 if (rbk // 4) * (cb.shape[1] // 4) * 8 <= 256 else 4)), 2)
    # whole-plane kernel (N <= 128): from 64 px it folds the rows of x into even / odd parts (half the multiply-adds of the two
    # large products)
    fold = 1 if N in (64, 128, 224) else 0          # 224 px: the half-plane row-blocked kernel folds too (288 px does not)
    _chk(lib().ee_oracle_hfs(_p(x), _p(y), _p(add), x.size // (N * N), N, r, _p(cb), _p(rb), _p(w), gamma, ks1, ks2, fold))
    return y


def add_square(x, stripe, table, eps, g=None):
    """Add_Square forward (g is None) or g * d(out)/d(x); stripe [B,C,W], table [n_sq, 2+C]."""
    x, stripe = _f32(x), _f32(stripe)
    B, C, H, W = x.shape
    table = _f32(table).reshape(-1, 2 + C) if table is not None and np.size(table) else np.zeros((0, 2 + C), np.float32)
    g = None if g is None else _f32(g)
    out = np.empty_like(x)
    lib().ee_oracle_add_square(_p(g), _p(x), _p(stripe), _p(table), _p(out), B, C, H, W, table.shape[0],
                               float(np.float32(eps)))
    return out


def _ew(name, *arrs, thr=None):
    arrs = [_f32(a) for a in arrs]
    out = np.empty_like(arrs[0])
    args = [_p(a) for a in arrs] + [_p(out), arrs[0].size]
    if thr is not None:
        args.append(float(np.float32(thr)))
    getattr(lib(), name)(*args)
    return out


def to_compare_fwd(v, thr): return _ew("ee_oracle_to_compare_fwd", v, thr=thr)
def to_compare_bwd(g, v, thr): return _ew("ee_oracle_to_compare_bwd", g, v, thr=thr)
def to_eq_fwd(v): return _ew("ee_oracle_to_eq_fwd", v)
def to_eq_bwd(g, v): return _ew("ee_oracle_to_eq_bwd", g, v)
def safe_sign_fwd(v): return _ew("ee_oracle_safe_sign_fwd", v)
def safe_sign_bwd(g, v): return _ew("ee_oracle_safe_sign_bwd", g, v)
