"""Torch-facing wrappers over the C ABI: raw ops and the autograd.Functions.

PyTorch is plumbing here (device memory, current stream, autograd bookkeeping); all arithmetic of
the hot path happens in libedge_b200.so.  Everything is CUDA-only and fp32-only: CPU tensors or
other dtypes raise -- there is no fallback path.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import EEParams, EE_VARIANT_BPDA, EE_VARIANT_CANNY, EE_VARIANT_STEP125

EE_LAYOUT_NCHW, EE_LAYOUT_NHWC = 0, 1

VARIANTS = {"step125": EE_VARIANT_STEP125, "canny": EE_VARIANT_CANNY, "bpda": EE_VARIANT_BPDA}

_SOBEL_X = np.array([[-0.5, 0.0, 0.5], [-1.0, 0.0, 1.0], [-0.5, 0.0, 0.5]], dtype=np.float32)


def make_params(variant, gauss, alpha=0.0, low=None, high=None, hysteresis=False, sobel=None, nan_compat=False):
    """Build an EEParams from module state (3x3 fp32 kernels) + forward() arguments.  nan_compat=True asks the backward
    for the reference's NaNs where the gradient magnitude is exactly 0 (EE_FLAG_NAN_COMPAT, include/edge_b200.h)."""
    p = EEParams()
    p.variant = VARIANTS[variant] if isinstance(variant, str) else int(variant)
    p.layout = _lib.EE_LAYOUT_NCHW
    g = np.asarray(gauss, dtype=np.float32).reshape(-1)
    s = (_SOBEL_X if sobel is None else np.asarray(sobel, dtype=np.float32)).reshape(-1)
    if g.size != 9 or s.size != 9:
        raise NotImplementedError("edge_b200 implements the 3x3 Gaussian / 3x3 Sobel of every reference config "
                                  "(k_gaussian=3, k_sobel=3); got %d / %d taps" % (g.size, s.size))
    for i in range(9):
        p.gauss[i] = float(g[i])
        p.sobel[i] = float(s[i])
    p.alpha = float(alpha)
    p.low_thr = 0.0 if low is None else float(low)
    p.high_thr = 0.0 if high is None else float(high)
    p.has_low = int(low is not None)
    p.has_high = int(high is not None)
    p.hysteresis = int(bool(hysteresis))
    p.flags = _lib.EE_FLAG_NAN_COMPAT if nan_compat else 0
    return p


def _chk(t, name, shape=None):
    return _chk_nocopy(t, name, shape).contiguous()


def _chk_nocopy(t, name, shape=None):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("edge_b200: %s is on %s; the B200 kernels are CUDA-only and there is no CPU fallback"
                           % (name, t.device))
    if t.dtype != torch.float32:
        raise TypeError("edge_b200: %s must be float32 (got %s)" % (name, t.dtype))
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError("edge_b200: %s has shape %s, expected %s" % (name, tuple(t.shape), tuple(shape)))
    return t


def _is_channels_last(t):
    return (t.dim() == 4 and t.shape[1] > 1 and not t.is_contiguous()
            and t.is_contiguous(memory_format=torch.channels_last))


def _nhwc_ok(x, params=None):
    """The fused kernels read torch.channels_last directly for C == 3 and W % 4 == 0 (include/edge_b200.h); the
    NaN-compatible backward is NCHW-only."""
    B, C, H, W = x.shape
    if params is not None and (params.flags & _lib.EE_FLAG_NAN_COMPAT):
        return False
    return _is_channels_last(x) and C == 3 and W % 4 == 0 and W >= 8 and H >= 4


def _as_layout(t, name, shape, nhwc):
    t = _chk_nocopy(t, name, shape)
    return t.contiguous(memory_format=torch.channels_last) if nhwc else t.contiguous()


def _rows_ok(t):
    """A non-dense 4-D view the library reads in place through ee_*_strided_f32: unit column stride (sliced batches,
    channel slices, expanded batches, spatial crops) and no negative strides."""
    return t.dim() == 4 and (t.shape[3] == 1 or t.stride(3) == 1) and min(t.stride()) >= 0


def _strides(t):
    return None if t is None or t.is_contiguous() else ctypes.byref(_lib.EEStrides(*t.stride()))


def _in_place_or_copy(t, name, shape=None):
    """The tensor itself when the library can read its strides, else a dense copy (never silently wrong)."""
    t = _chk_nocopy(t, name, shape)
    return t if (t.is_contiguous() or _rows_ok(t)) else t.contiguous()


def _with_layout(params, nhwc):
    want = EE_LAYOUT_NHWC if nhwc else EE_LAYOUT_NCHW
    if params.layout == want:
        return params
    q = EEParams.from_buffer_copy(params)
    q.layout = want
    return q


def _out_like(x, out):
    if out is None:
        return torch.empty_like(x)              # preserves x's memory format
    if not (out.is_cuda and out.dtype == torch.float32 and out.shape == x.shape and out.stride() == x.stride()):
        raise ValueError("edge_b200: `out` must be a dense float32 CUDA tensor with the shape and strides of the input")
    return out


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream(t):
    """cudaStream_t of the CURRENT stream of t's device (what autograd / DataParallel threads expect), as a plain int:
    ctypes converts it to void* itself, which is cheaper than building a c_void_p per call."""
    if _raw_stream is not None:
        return _raw_stream(t.device.index)
    return torch.cuda.current_stream(t.device).cuda_stream


class _on_device:
    """`with torch.cuda.device(t.device)` costs ~5 us per call; skip it when t's device is already current."""
    __slots__ = ("guard",)

    def __init__(self, t):
        self.guard = None if t.device.index == torch.cuda.current_device() else torch.cuda.device(t.device)

    def __enter__(self):
        if self.guard is not None:
            self.guard.__enter__()

    def __exit__(self, *exc):
        if self.guard is not None:
            self.guard.__exit__(*exc)


def _ptr(t):
    return None if t is None else t.data_ptr()


# --------------------------------------------------------------------------------------------
# raw ops (no autograd)
# --------------------------------------------------------------------------------------------
def edge_map(x, params):
    """edge[B,1,H,W] = filter(x[B,C,H,W]) -- ee_edge_fwd_f32."""
    x = _in_place_or_copy(x, "img")
    if x.dim() != 4:
        raise ValueError("img must be [B,C,H,W]")
    B, C, H, W = x.shape
    edge = torch.empty((B, 1, H, W), dtype=torch.float32, device=x.device)
    if x.numel() == 0:
        return edge
    with _on_device(x):
        if x.is_contiguous():
            rc = _lib.load().ee_edge_fwd_f32(_ptr(x), _ptr(edge), B, C, H, W, ctypes.byref(params), _stream(x))
        else:                       # a view with unit column stride: read in place
            rc = _lib.load().ee_edge_fwd_strided_f32(_ptr(x), _strides(x), _ptr(edge), None, B, C, H, W, ctypes.byref(params), _stream(x))
    _lib.check(rc, "ee_edge_fwd_f32")
    return edge


def edge_map_backward(g_edge, x, params):
    x = _in_place_or_copy(x, "img")
    B, C, H, W = x.shape
    g_edge = _in_place_or_copy(g_edge, "grad_edge", (B, 1, H, W))
    g_x = torch.empty((B, C, H, W), dtype=torch.float32, device=x.device)
    if x.numel() == 0:
        return g_x
    with _on_device(x):
        if x.is_contiguous() and g_edge.is_contiguous():
            rc = _lib.load().ee_edge_bwd_f32(_ptr(g_edge), _ptr(x), _ptr(g_x), B, C, H, W, ctypes.byref(params), _stream(x))
        else:
            rc = _lib.load().ee_edge_bwd_strided_f32(_ptr(g_edge), _strides(g_edge), _ptr(x), _strides(x), _ptr(g_x), None,
                                                     B, C, H, W, ctypes.byref(params), _stream(x))
    _lib.check(rc, "ee_edge_bwd_f32")
    return g_x


def edge_blend(x, base, params, w, want_edge=False, out=None):
    """out = clamp(base + w*filter(x), 0, 1) in one pass -- ee_edge_blend_fwd_f32.
    `out` (optional) is a preallocated result buffer; it must not alias x or base.  channels_last
    (NHWC) inputs with C == 3 are read in place; other layouts are made NCHW-contiguous first."""
    x = _chk_nocopy(x, "img")
    if x.dim() != 4:
        raise ValueError("img must be [B,C,H,W]")
    nhwc = _nhwc_ok(x)
    if not nhwc and not (x.is_contiguous() and base.is_contiguous()):
        return _edge_blend_strided(x, base, params, w, want_edge, out)
    x = _as_layout(x, "img", None, nhwc)
    B, C, H, W = x.shape
    base = _as_layout(base, "base", x.shape, nhwc)
    out = _out_like(x, out)
    edge = torch.empty((B, 1, H, W), dtype=torch.float32, device=x.device) if want_edge else None
    if x.numel():
        p = _with_layout(params, nhwc)
        with _on_device(x):
            rc = _lib.load().ee_edge_blend_fwd_f32(_ptr(x), _ptr(base), _ptr(out), _ptr(edge), B, C, H, W,
                                                   ctypes.byref(p), float(w), _stream(x))
        _lib.check(rc, "ee_edge_blend_fwd_f32")
    return (out, edge) if want_edge else out


def _edge_blend_strided(x, base, params, w, want_edge, out):
    """edge_blend on views the library reads in place (unit column stride): no .contiguous() pass over HBM."""
    x = _in_place_or_copy(x, "img")
    B, C, H, W = x.shape
    base = _in_place_or_copy(base, "base", x.shape)
    if out is None:
        out = torch.empty((B, C, H, W), dtype=torch.float32, device=x.device)
    elif not (out.is_cuda and out.dtype == torch.float32 and out.shape == x.shape and (out.is_contiguous() or _rows_ok(out))):
        raise ValueError("edge_b200: `out` must be a float32 CUDA tensor of the input's shape with unit column stride")
    edge = torch.empty((B, 1, H, W), dtype=torch.float32, device=x.device) if want_edge else None
    if x.numel():
        p = _with_layout(params, False)
        with _on_device(x):
            rc = _lib.load().ee_edge_blend_fwd_strided_f32(_ptr(x), _strides(x), _ptr(base), _strides(base), _ptr(out), _strides(out),
                                                           _ptr(edge), None, B, C, H, W, ctypes.byref(p), float(w), _stream(x))
        _lib.check(rc, "ee_edge_blend_fwd_strided_f32")
    return (out, edge) if want_edge else out


def edge_blend_backward(g_out, x, base, params, w, need_x=True, need_base=True, g_x=None, g_base=None):
    """(g_x, g_base) of edge_blend in one pass -- ee_edge_blend_bwd_f32.  g_x / g_base may be
    preallocated buffers (not aliasing any input)."""
    x = _chk_nocopy(x, "img")
    nhwc = _nhwc_ok(x, params)
    if not nhwc and not _is_channels_last(x) and not (x.is_contiguous() and base.is_contiguous() and g_out.is_contiguous()):
        return _edge_blend_backward_strided(g_out, x, base, params, w, need_x, need_base, g_x, g_base)
    x = _as_layout(x, "img", None, nhwc)
    B, C, H, W = x.shape
    base = _as_layout(base, "base", x.shape, nhwc)
    g_out = _as_layout(g_out, "grad_out", x.shape, nhwc)
    g_x = _out_like(x, g_x) if need_x else None
    g_base = _out_like(x, g_base) if need_base else None
    if x.numel() and (need_x or need_base):
        p = _with_layout(params, nhwc)
        with _on_device(x):
            rc = _lib.load().ee_edge_blend_bwd_f32(_ptr(g_out), _ptr(x), _ptr(base), _ptr(g_x), _ptr(g_base),
                                                   B, C, H, W, ctypes.byref(p), float(w), _stream(x))
        _lib.check(rc, "ee_edge_blend_bwd_f32")
    return g_x, g_base


def pgd_iteration(x, base, g_out, x0, params, w, alpha_signed, eps, out=None, g_x=None, g_base=None, x_next=None,
                  want_out=True, want_base=True):
    """One iteration of the edge-enhanced PGD hot path in ONE C-ABI call (ee_edge_pgd_iteration_f32): fused forward
    (skipped with want_out=False), fused backward of `g_out`, sign / project / clamp update from the edge-path gradient.
    Dense NCHW tensors; returns (out, g_x, g_base, x_next)."""
    x, base, g_out, x0 = _same(x, base, g_out, x0)
    if x.dim() != 4:
        raise ValueError("img must be [B,C,H,W]")
    B, C, H, W = x.shape
    out = _out_like(x, out) if want_out else None
    g_x = _out_like(x, g_x)
    g_base = _out_like(x, g_base) if want_base else None
    x_next = _out_like(x, x_next)
    if x.numel():
        p = _with_layout(params, False)
        with _on_device(x):
            rc = _lib.load().ee_edge_pgd_iteration_f32(_ptr(x), _ptr(base), _ptr(g_out), _ptr(x0), _ptr(out), _ptr(g_x),
                                                       _ptr(g_base), _ptr(x_next), B, C, H, W, ctypes.byref(p), float(w),
                                                       float(alpha_signed), float(eps), _stream(x))
        _lib.check(rc, "ee_edge_pgd_iteration_f32")
    return out, g_x, g_base, x_next


def _gauss9(gauss):
    g = np.ascontiguousarray(np.asarray(gauss, dtype=np.float32).reshape(-1))
    if g.size != 9:
        raise NotImplementedError("edge_b200: the gf option uses the 3x3 Gaussian of the reference (resnet_EE.py:133-136)")
    return (ctypes.c_float * 9)(*[float(v) for v in g])


def gf_blend(edge, base, gauss, w, out=None):
    """out = clamp(base + w * conv2d(edge, gauss, padding=1), 0, 1): the with_gf=True blend (resnet_EE.py:185-191)."""
    base = _chk(base, "base")
    B, C, H, W = base.shape
    edge = _chk(edge, "edge", (B, 1, H, W))
    out = _out_like(base, out)
    if base.numel():
        with _on_device(base):
            rc = _lib.load().ee_gf_blend_fwd_f32(_ptr(edge), _ptr(base), _ptr(out), B, C, H, W, _gauss9(gauss), float(w),
                                                 _stream(base))
        _lib.check(rc, "ee_gf_blend_fwd_f32")
    return out


def gf_blend_backward(g_out, edge, base, gauss, w, need_edge=True, need_base=True):
    """(g_edge, g_base) of gf_blend in one pass."""
    base = _chk(base, "base")
    B, C, H, W = base.shape
    edge = _chk(edge, "edge", (B, 1, H, W))
    g_out = _chk(g_out, "grad_out", base.shape)
    g_edge = torch.empty_like(edge) if need_edge else None
    g_base = torch.empty_like(base) if need_base else None
    if base.numel() and (need_edge or need_base):
        with _on_device(base):
            rc = _lib.load().ee_gf_blend_bwd_f32(_ptr(g_out), _ptr(edge), _ptr(base), _ptr(g_edge), _ptr(g_base), B, C, H, W,
                                                 _gauss9(gauss), float(w), _stream(base))
        _lib.check(rc, "ee_gf_blend_bwd_f32")
    return g_edge, g_base


def _edge_blend_backward_strided(g_out, x, base, params, w, need_x, need_base, g_x, g_base):
    x = _in_place_or_copy(x, "img")
    B, C, H, W = x.shape
    base = _in_place_or_copy(base, "base", x.shape)
    g_out = _in_place_or_copy(g_out, "grad_out", x.shape)

    def result(buf):
        if buf is None:
            return torch.empty((B, C, H, W), dtype=torch.float32, device=x.device)
        if not (buf.is_cuda and buf.dtype == torch.float32 and buf.shape == x.shape and (buf.is_contiguous() or _rows_ok(buf))):
            raise ValueError("edge_b200: gradient buffers must be float32 CUDA tensors of the input's shape with unit column stride")
        return buf
    g_x = result(g_x) if need_x else None
    g_base = result(g_base) if need_base else None
    if x.numel() and (need_x or need_base):
        p = _with_layout(params, False)
        with _on_device(x):
            rc = _lib.load().ee_edge_blend_bwd_strided_f32(_ptr(g_out), _strides(g_out), _ptr(x), _strides(x), _ptr(base), _strides(base),
                                                           _ptr(g_x), _strides(g_x), _ptr(g_base), _strides(g_base), B, C, H, W,
                                                           ctypes.byref(p), float(w), _stream(x))
        _lib.check(rc, "ee_edge_blend_bwd_strided_f32")
    return g_x, g_base


def _same(*ts):
    ref = ts[0]
    out = []
    for i, t in enumerate(ts):
        t = _chk(t, "operand %d" % i)
        if t.shape != ref.shape or t.device != ref.device:
            raise ValueError("edge_b200: operand %d has shape/device %s/%s, expected %s/%s"
                             % (i, tuple(t.shape), t.device, tuple(ref.shape), ref.device))
        out.append(t)
    return out


def pgd_linf_step(x, grad, x0, alpha_signed, eps, lo=0.0, hi=1.0, out=None):
    """clamp(min(max(x + alpha_signed*sign(grad), x0-eps), x0+eps), lo, hi); `out` may be x."""
    x, grad, x0 = _same(x, grad, x0)
    out = _out_like(x, out)
    if x.numel():
        with _on_device(x):
            rc = _lib.load().ee_pgd_linf_step_f32(_ptr(x), _ptr(grad), _ptr(x0), _ptr(out), x.numel(),
                                                  float(alpha_signed), float(eps), float(lo), float(hi), _stream(x))
        _lib.check(rc, "ee_pgd_linf_step_f32")
    return out


def fgsm_step(x, grad, alpha_signed, lo=0.0, hi=1.0, out=None):
    x, grad = _same(x, grad)
    out = _out_like(x, out)
    if x.numel():
        with _on_device(x):
            rc = _lib.load().ee_fgsm_step_f32(_ptr(x), _ptr(grad), _ptr(out), x.numel(), float(alpha_signed),
                                              float(lo), float(hi), _stream(x))
        _lib.check(rc, "ee_fgsm_step_f32")
    return out


def free_at_step_(delta, grad, x0=None, alpha=0.0, eps=0.0, lo=0.0, hi=1.0, want_adv=True):
    """In place: delta = clamp(delta + alpha*sign(grad), -eps, eps); returns x_adv = clamp(x0 + delta, lo, hi)
    (or None when want_adv is False).  `delta` must be contiguous (it is updated in place)."""
    if not (isinstance(delta, torch.Tensor) and delta.is_cuda and delta.dtype == torch.float32 and delta.is_contiguous()):
        raise ValueError("edge_b200: delta must be a contiguous float32 CUDA tensor (updated in place)")
    grad = _chk(grad, "grad", delta.shape)
    x_adv = None
    if want_adv:
        x0 = _chk(x0, "x0", delta.shape)
        x_adv = torch.empty_like(delta)
    if delta.numel():
        with _on_device(delta):
            rc = _lib.load().ee_free_at_step_f32(_ptr(delta), _ptr(grad), _ptr(x0 if want_adv else None), _ptr(x_adv),
                                                 delta.numel(), float(alpha), float(eps), float(lo), float(hi),
                                                 _stream(delta))
        _lib.check(rc, "ee_free_at_step_f32")
    return x_adv


def add_clamp(x, noise, lo=0.0, hi=1.0, out=None):
    """clamp(x + noise, lo, hi): the random start of the attacks (utils/attacks.py:15-17) in one pass."""
    x, noise = _same(x, noise)
    out = _out_like(x, out)
    if x.numel():
        with _on_device(x):
            rc = _lib.load().ee_add_clamp_f32(_ptr(x), _ptr(noise), _ptr(out), x.numel(), float(lo), float(hi), _stream(x))
        _lib.check(rc, "ee_add_clamp_f32")
    return out


def avmixup_mix(x_adv, inputs, weight, gamma):
    """AVmixup vertex + mix (utils/attacks.py:469-471, :476): weight is the float64 per-sample tensor ([B] or [B,1,1,1])."""
    x_adv, inputs = _same(x_adv, inputs)
    B = x_adv.shape[0]
    if not (isinstance(weight, torch.Tensor) and weight.is_cuda and weight.dtype == torch.float64 and weight.numel() == B):
        raise ValueError("edge_b200: avmixup weight must be a float64 CUDA tensor with one entry per sample")
    weight = weight.reshape(B).contiguous()
    out = torch.empty_like(x_adv)
    if x_adv.numel():
        with _on_device(x_adv):
            rc = _lib.load().ee_avmixup_mix_f32(_ptr(x_adv), _ptr(inputs), _ptr(weight), _ptr(out), B,
                                                x_adv.numel() // B, float(gamma), _stream(x_adv))
        _lib.check(rc, "ee_avmixup_mix_f32")
    return out


def cw_linf_step(adv, grad, x, min_x, max_x, step, magnitude, out=None):
    adv, grad, x, min_x, max_x = _same(adv, grad, x, min_x, max_x)
    out = _out_like(adv, out)
    if adv.numel():
        with _on_device(adv):
            rc = _lib.load().ee_cw_linf_step_f32(_ptr(adv), _ptr(grad), _ptr(x), _ptr(min_x), _ptr(max_x), _ptr(out),
                                                 adv.numel(), float(step), float(magnitude), _stream(adv))
        _lib.check(rc, "ee_cw_linf_step_f32")
    return out


def pgd_l2_step(x, grad, x0, step, eps):
    """TRADES PGD-L2 update with per-sample RMS norms (first dim = batch)."""
    x, grad, x0 = _same(x, grad, x0)
    out = torch.empty_like(x)
    if x.numel():
        B = x.shape[0]
        with _on_device(x):
            rc = _lib.load().ee_pgd_l2_step_f32(_ptr(x), _ptr(grad), _ptr(x0), _ptr(out), B, x.numel() // B,
                                                float(step), float(eps), _stream(x))
        _lib.check(rc, "ee_pgd_l2_step_f32")
    return out


_HFS_TABLES = {}


def hfs_tables(N, r, device):
    """Device copies of the real Fourier bases / mixing weights of HighFreqSuppress(N, N, r) (cached per device)."""
    key = (N, r, str(device))
    t = _HFS_TABLES.get(key)
    if t is None:
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
            cb[N // 2, r:] = 0.0       # sin(k pi) is exactly 0: the folded kernel's sine chains start at w = N/2
        alpha = np.zeros(NIp); beta = np.zeros(NJp)
        alpha[0] = 1.0 / N; alpha[1:r] = 2.0 / N; alpha[r] = 1.0 / N; alpha[r + 1:2 * r] = 2.0 / N; alpha[2 * r] = 1.0 / N
        beta[0] = 1.0 / N; beta[1:NJ] = 2.0 / N
        w = np.outer(alpha, beta)
        t = tuple(torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(device) for a in (cb, rb, w)) + (float(np.float32(2.0 / (N * N))),)
        _HFS_TABLES[key] = t
    return t


def hfs_supported(N, r, impl='native'):
    L = _lib.load()
    return bool((L.ee_hfs_tc_supported if impl == 'tcgen05' else L.ee_hfs_supported)(int(N), int(r)))


def hfs(x, r, out=None, add=None, impl='native'):
    """y = HighFreqSuppress(N, N, r)(x) [+ add] for [..., N, N] planes -- ee_hfs_f32 (one kernel; self-adjoint, so the same
    call on the upstream gradient is the backward; `add` lets it accumulate into another gradient, and may be `out`).
    impl='tcgen05' selects ee_hfs_tc_f32, the tensor-core variant (64 x 64 / r 8 only; 3 x TF32, not bit-identical)."""
    if impl not in ('native', 'tcgen05'):
        raise ValueError("edge_b200: hfs impl must be 'native' or 'tcgen05'")
    x = _chk(x, "x")
    if add is not None:
        add = _chk(add, "add", x.shape)
    N = x.shape[-1]
    if x.dim() < 2 or x.shape[-2] != N:
        raise ValueError("edge_b200: hfs needs square [..., N, N] planes")
    out = _out_like(x, out)
    if x.numel():
        cb, rb, w, gamma = hfs_tables(N, r, x.device)
        with _on_device(x):
            fn = _lib.load().ee_hfs_tc_f32 if impl == 'tcgen05' else _lib.load().ee_hfs_f32
            rc = fn(_ptr(x), _ptr(out), _ptr(add), x.numel() // (N * N), N, int(r), _ptr(cb), _ptr(rb), _ptr(w), gamma, _stream(x))
        _lib.check(rc, "ee_hfs_tc_f32" if impl == 'tcgen05' else "ee_hfs_f32")
    return out


def _square_operands(x, stripe, table):
    x = _chk(x, "x")
    if x.dim() != 4:
        raise ValueError("edge_b200: add_square needs a [B,C,H,W] tensor")
    B, C, H, W = x.shape
    stripe = _chk(stripe, "stripe", (B, C, W))
    n_sq = 0
    if table is not None and table.numel():
        table = _chk(table, "table")
        if table.dim() != 2 or table.shape[1] != 2 + C:
            raise ValueError("edge_b200: square table must be [n_queries, 2 + C]")
        n_sq = table.shape[0]
    else:
        table = None
    return x, stripe, table, n_sq


def add_square(x, stripe, table, eps, out=None):
    """Add_Square.forward (utils/core.py:640-655) from pre-drawn random numbers: stripe[B,C,W] = signs of the
    column stripes, table[n_queries, 2+C] = (vh, s, 2*eps*sign_c ...) per query.  One pass."""
    x, stripe, table, n_sq = _square_operands(x, stripe, table)
    out = _out_like(x, out)
    B, C, H, W = x.shape
    if x.numel():
        with _on_device(x):
            rc = _lib.load().ee_add_square_fwd_f32(_ptr(x), _ptr(stripe), _ptr(table), _ptr(out), B, C, H, W, n_sq,
                                                   float(eps), _stream(x))
        _lib.check(rc, "ee_add_square_fwd_f32")
    return out


def add_square_backward(g, x, stripe, table, eps, out=None):
    """g_x = g * d(add_square)/d(x): autograd through the clamps and the two-sided projection, recomputed from x."""
    x, stripe, table, n_sq = _square_operands(x, stripe, table)
    g = _chk(g, "g", x.shape)
    out = _out_like(x, out)
    B, C, H, W = x.shape
    if x.numel():
        with _on_device(x):
            rc = _lib.load().ee_add_square_bwd_f32(_ptr(g), _ptr(x), _ptr(stripe), _ptr(table), _ptr(out), B, C, H, W,
                                                   n_sq, float(eps), _stream(x))
        _lib.check(rc, "ee_add_square_bwd_f32")
    return out


def _ew(name, args, n, ref, thr=None):
    out = torch.empty_like(ref)
    if n:
        call = [_ptr(a) for a in args] + [_ptr(out), n]
        if thr is not None:
            call.append(float(thr))
        call.append(_stream(ref))
        with _on_device(ref):
            rc = getattr(_lib.load(), name)(*call)
        _lib.check(rc, name)
    return out


# --------------------------------------------------------------------------------------------
# autograd Functions
# --------------------------------------------------------------------------------------------
class EdgeMapFn(torch.autograd.Function):
    """edge = CannyFilter*(img); backward recomputes the forward intermediates from img."""

    @staticmethod
    def forward(ctx, img, params):
        ctx.params = params
        ctx.save_for_backward(img)
        return edge_map(img, params)

    @staticmethod
    def backward(ctx, g_edge):
        (img,) = ctx.saved_tensors
        g = edge_map_backward(g_edge, img, ctx.params) if ctx.needs_input_grad[0] else None
        return g, None


class EdgeEnhanceFn(torch.autograd.Function):
    """out = clamp(base + w * CannyFilter*(img), 0, 1), one HBM pass per direction."""

    @staticmethod
    def forward(ctx, img, base, params, w):
        ctx.params = params
        ctx.w = float(w)
        ctx.save_for_backward(img, base)
        return edge_blend(img, base, params, w)

    @staticmethod
    def backward(ctx, g_out):
        img, base = ctx.saved_tensors
        need_x, need_base = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        g_x, g_base = edge_blend_backward(g_out, img, base, ctx.params, ctx.w, need_x, need_base)
        return g_x, g_base, None, None


class GfBlendFn(torch.autograd.Function):
    """out = clamp(base + w * gauss3x3_zero_pad(edge), 0, 1) -- the `with_gf=True` blend of the *_EE models."""

    @staticmethod
    def forward(ctx, edge, base, gauss, w):
        ctx.gauss, ctx.w = gauss, float(w)
        ctx.save_for_backward(edge, base)
        return gf_blend(edge, base, gauss, w)

    @staticmethod
    def backward(ctx, g_out):
        edge, base = ctx.saved_tensors
        g_edge, g_base = gf_blend_backward(g_out, edge, base, ctx.gauss, ctx.w, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return g_edge, g_base, None, None


class HfsFn(torch.autograd.Function):
    """x_hfs = HighFreqSuppress(x); the operator is symmetric, so the backward applies it to the upstream gradient."""

    @staticmethod
    def forward(ctx, x, r, impl='native'):
        ctx.r, ctx.impl = r, impl
        return hfs(x, r, impl=impl)

    @staticmethod
    def backward(ctx, g):
        return hfs(g, ctx.r, impl=ctx.impl), None, None


class EdgeEnhanceFrontFn(torch.autograd.Function):
    """The whole front end of the *_EE models (resnet_EE.py:176-191) as three kernels per direction:
        forward : base = HFS(x) ; out = clamp(base + w * edge(x), 0, 1)
        backward: (g_x, g_base) = edge_blend_backward(g) ; g_x = HFS(g_base) + g_x     (HFS is symmetric; the sum is fused
                  into the low-pass kernel's store, so autograd never runs a separate accumulation pass)"""

    @staticmethod
    def forward(ctx, x, r, params, w, hfs_impl='native'):
        base = hfs(x, r, impl=hfs_impl)
        ctx.r, ctx.params, ctx.w, ctx.hfs_impl = r, params, w, hfs_impl
        ctx.save_for_backward(x, base)
        return edge_blend(x, base, params, w)

    @staticmethod
    def backward(ctx, g):
        x, base = ctx.saved_tensors
        g_x, g_base = edge_blend_backward(g, x, base, ctx.params, ctx.w)
        return hfs(g_base, ctx.r, out=g_x, add=g_x, impl=ctx.hfs_impl), None, None, None, None


class AddSquareFn(torch.autograd.Function):
    """x_square = Add_Square(x) with the random draws passed in; backward recomputes the multiplier from x."""

    @staticmethod
    def forward(ctx, x, stripe, table, eps):
        ctx.eps = eps
        ctx.save_for_backward(x, stripe, table)
        return add_square(x, stripe, table, eps)

    @staticmethod
    def backward(ctx, g):
        x, stripe, table = ctx.saved_tensors
        return add_square_backward(g, x, stripe, table, ctx.eps), None, None, None


class ToCompareFn(torch.autograd.Function):
    """utils/core.py:329-358."""

    @staticmethod
    def forward(ctx, input, threshold):
        thr = float(threshold)
        ctx.thr = thr
        x = _chk(input, "input")
        ctx.save_for_backward(x)
        return _ew("ee_to_compare_fwd_f32", [x], x.numel(), x, thr)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        g = _chk(g, "grad", x.shape)
        return _ew("ee_to_compare_bwd_f32", [g, x], x.numel(), x, ctx.thr), None


class ToEqFn(torch.autograd.Function):
    """utils/core.py:361-382."""

    @staticmethod
    def forward(ctx, input):
        x = _chk(input, "input")
        ctx.save_for_backward(x)
        return _ew("ee_to_eq_fwd_f32", [x], x.numel(), x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        g = _chk(g, "grad", x.shape)
        return _ew("ee_to_eq_bwd_f32", [g, x], x.numel(), x)


class SafeSignFn(torch.autograd.Function):
    """BinaryConnectDeterministic, utils/core.py:121-145."""

    @staticmethod
    def forward(ctx, input):
        x = _chk(input, "input")
        ctx.save_for_backward(x)
        return _ew("ee_safe_sign_fwd_f32", [x], x.numel(), x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        g = _chk(g, "grad", x.shape)
        return _ew("ee_safe_sign_bwd_f32", [g, x], x.numel(), x)


def safe_sign(t):
    """safeSign of utils/core.py:115-118 (no autograd)."""
    x = _chk(t, "tensor")
    return _ew("ee_safe_sign_fwd_f32", [x], x.numel(), x)
