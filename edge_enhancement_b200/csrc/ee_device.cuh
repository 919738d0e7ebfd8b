// ee_device.cuh -- device-side building blocks of the edge-enhancement kernels (sm_100a).
//
// Canonical arithmetic: every fp32 expression below is written with explicit
// fmaf() where a fused multiply-add is intended and the translation unit is compiled with
// -fmad=false, so ptxas never contracts a*b+c on its own; '/' and sqrtf are the IEEE-rounded
// versions (-prec-div=true -prec-sqrt=true, no -use_fast_math, denormals kept).  The edge
// mask is a threshold on these values, so the evaluation order IS the specification.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ee {

// ---------------------------------------------------------------------------------------------
// vector <-> array helpers.  VEC is 4 (128-bit path, W % 4 == 0 and 16-byte aligned tensors) or
// 1 (any shape / alignment).
// ---------------------------------------------------------------------------------------------
template <int VEC>
__device__ __forceinline__ void ld_vec(const float* __restrict__ p, float (&v)[VEC]) {
    if constexpr (VEC == 4) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
        v[0] = *p;
    }
}
template <int VEC>
__device__ __forceinline__ void st_vec(float* __restrict__ p, const float (&v)[VEC]) {
    if constexpr (VEC == 4) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
        *p = v[0];
    }
}
// read-only global load through the non-coherent path (inputs are never written by the kernel)
template <int VEC>
__device__ __forceinline__ void ldg_vec(const float* __restrict__ p, float (&v)[VEC]) {
    if constexpr (VEC == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
        v[0] = __ldg(p);
    }
}
// streaming global store (outputs are not re-read by this kernel: evict-first)
template <int VEC>
__device__ __forceinline__ void stg_vec(float* __restrict__ p, const float (&v)[VEC]) {
    if constexpr (VEC == 4) {
        __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
    } else {
        __stcs(p, v[0]);
    }
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// ---------------------------------------------------------------------------------------------
// A "window" is VEC consecutive pixels of one smem plane row plus the left and right
// neighbour: e[0] | e[1..VEC] | e[VEC+1].  Replicate (clamped column) or zero extension.
// rowp == nullptr means "row outside the image" in the zero-extended (adjoint) case.
// ---------------------------------------------------------------------------------------------
template <int VEC, bool ZERO_EXT>
__device__ __forceinline__ void load_win(const float* rowp, int col, int W, float (&e)[VEC + 2]) {
    if (ZERO_EXT && rowp == nullptr) {
#pragma unroll
        for (int k = 0; k < VEC + 2; ++k) e[k] = 0.0f;
        return;
    }
    float c[VEC];
    ld_vec<VEC>(rowp + col, c);
#pragma unroll
    for (int k = 0; k < VEC; ++k) e[k + 1] = c[k];
    const int cl = col - 1, cr = col + VEC;
    if (ZERO_EXT) {
        e[0] = (cl >= 0) ? rowp[cl] : 0.0f;
        e[VEC + 1] = (cr < W) ? rowp[cr] : 0.0f;
    } else {
        e[0] = rowp[cl >= 0 ? cl : 0];
        e[VEC + 1] = rowp[cr < W ? cr : W - 1];
    }
}
// single-column zero-extended window at an arbitrary column q in [-1, W] (ring columns)
__device__ __forceinline__ void load_win1_zero(const float* rowp, int q, int W, float (&e)[3]) {
    if (rowp == nullptr) { e[0] = e[1] = e[2] = 0.0f; return; }
    e[0] = (q - 1 >= 0 && q - 1 < W) ? rowp[q - 1] : 0.0f;
    e[1] = (q >= 0 && q < W) ? rowp[q] : 0.0f;
    e[2] = (q + 1 >= 0 && q + 1 < W) ? rowp[q + 1] : 0.0f;
}

// ---------------------------------------------------------------------------------------------
// 3x3 symmetric Gaussian over three windows (rows above / centre / below):
//     e = l + r;  P = fma(c1, m, c0*e) (outer rows);  Q = fma(c2, m, c1*e) (centre row)
//     out = (P_up + Q_mid) + P_dn
// Used by the forward blur (replicate windows, utils/core.py:560-563) and by its adjoint (zero
// windows; the kernel is symmetric so corr^T has the same taps).
// ---------------------------------------------------------------------------------------------
template <int VEC>
__device__ __forceinline__ void gauss3(const float (&u)[VEC + 2], const float (&m)[VEC + 2],
                                       const float (&d)[VEC + 2], float c0, float c1, float c2,
                                       float (&out)[VEC]) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        const float pu = fmaf(c1, u[k + 1], c0 * (u[k] + u[k + 2]));
        const float qm = fmaf(c2, m[k + 1], c1 * (m[k] + m[k + 2]));
        const float pd = fmaf(c1, d[k + 1], c0 * (d[k] + d[k + 2]));
        out[k] = (pu + qm) + pd;
    }
}

// ---------------------------------------------------------------------------------------------
// Sobel x / y over three replicate windows of the blurred plane, then the true division by C
// (utils/core.py:565-570):  D = r - l ; V = fma(.5, l + r, m)
//     Sgx = fma(.5, D_up + D_dn, D_mid) ; Sgy = V_dn - V_up
// ---------------------------------------------------------------------------------------------
template <int VEC>
__device__ __forceinline__ void sobel3(const float (&u)[VEC + 2], const float (&m)[VEC + 2],
                                       const float (&d)[VEC + 2], float fC, float (&gx1)[VEC],
                                       float (&gy1)[VEC]) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        const float du = u[k + 2] - u[k], dm = m[k + 2] - m[k], dd = d[k + 2] - d[k];
        const float vu = fmaf(0.5f, u[k] + u[k + 2], u[k + 1]);
        const float vd = fmaf(0.5f, d[k] + d[k + 2], d[k + 1]);
        const float sgx = fmaf(0.5f, du + dd, dm);
        const float sgy = vd - vu;
        gx1[k] = sgx / fC;
        gy1[k] = sgy / fC;
    }
}

// Adjoint of the Sobel pair into the padded frame, from zero-extended windows of a = dL/dSgx
// and b = dL/dSgy (rows p-1, p, p+1):
//     HA = a(q-1) - a(q+1) ; HB = fma(.5, b(q-1) + b(q+1), b(q))
//     T  = fma(.5, HA_up + HA_dn, HA_mid) + (HB_up - HB_dn)
template <int N>
__device__ __forceinline__ void sobel3_adj(const float (&au)[N + 2], const float (&am)[N + 2],
                                           const float (&ad)[N + 2], const float (&bu)[N + 2],
                                           const float (&bm)[N + 2], const float (&bd)[N + 2],
                                           float (&out)[N]) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const float hau = au[k] - au[k + 2], ham = am[k] - am[k + 2], had = ad[k] - ad[k + 2];
        const float hbu = fmaf(0.5f, bu[k] + bu[k + 2], bu[k + 1]);
        const float hbd = fmaf(0.5f, bd[k] + bd[k + 2], bd[k + 1]);
        (void)bm;
        const float xa = fmaf(0.5f, hau + had, ham);
        const float yb = hbu - hbd;
        out[k] = xa + yb;
    }
}

// magnitude exactly as (gx^2 + gy^2) ** 0.5 evaluates in torch (x*x, +, sqrt: three roundings)
__device__ __forceinline__ float magnitude(float gx1, float gy1) {
    return sqrtf(gx1 * gx1 + gy1 * gy1);
}

// To_compare.forward (utils/core.py:338-347): two masked writes in sequence; with a negative
// threshold the zeros written first become ones (0 > thr).  NaN stays NaN.
__device__ __forceinline__ float to_compare(float v, float thr) {
    if (v > thr) return 1.0f;
    if (v <= thr) return (0.0f > thr) ? 1.0f : 0.0f;
    return v;
}
// (safeSign(v - thr) + 1) / 2, utils/core.py:115-118,:299-310
__device__ __forceinline__ float sign_step(float v, float thr) { return (v - thr > 0.0f) ? 1.0f : 0.0f; }
// backward windows: To_compare.backward (core.py:350-358), BinaryConnectDeterministic.backward (:138-145)
// (masked assignments grad_input[mask] = 0, i.e. a select: a non-finite gradient outside the window is 0)
__device__ __forceinline__ float ste_sel(float g, float v, float thr) { return (v > thr && v <= 1.001f) ? g : 0.0f; }
__device__ __forceinline__ float bcd_sel(float g, float v, float thr) { return (fabsf(v - thr) > 1.001f) ? 0.0f : g; }

// torch.clamp(v, 0, 1): NaN propagates (fminf/fmaxf would drop it)
__device__ __forceinline__ float clamp01_nan(float v) {
    return (v != v) ? v : fminf(fmaxf(v, 0.0f), 1.0f);
}

// Direction pair used by non-maximum suppression (utils/core.py:258-260,:270,:275-281), evaluated without
// atan or division: bin boundaries are |gy/gx| = t_i = tan((2i+1)pi/16); with a = |gx|, s = sign(gx)*gy,
// m = #{i : |s| > t_i*a} and dir = (s > 0) ? m mod 4 : (4 - m) mod 4   (see oracle orient_dir).
__device__ __forceinline__ int orient_dir(float gx1, float gy1) {
    const float a = fabsf(gx1);
    const float s = (gx1 < 0.0f) ? -gy1 : gy1;
    const float as = fabsf(s);
    int m = 0;
    m += (as > 0.19891236737965800691f * a);
    m += (as > 0.66817863791929891999f * a);
    m += (as > 1.49660576266548901760f * a);
    m += (as > 5.02733949212584810451f * a);
    return (s > 0.0f) ? (m & 3) : ((4 - m) & 3);
}

// dL/d(mag) -> (dL/dSgx, dL/dSgy):  mag = u^.5, u = gx1^2 + gy1^2 ; autograd evaluates
// g*0.5*u^-.5, then *2*gx1, then /C (three divisions); canonical form: (g / (mag*C)) * gx1, one
// division, equal up to ~2 ulp.  Sub-gradient at mag == 0 is 0 -- unless nan_compat asks for what autograd does
// there: g * 0.5 * 0^-0.5 * 2 * 0 = NaN whatever g is (EE_FLAG_NAN_COMPAT, include/edge_b200.h).
__device__ __forceinline__ void mag_backward(float gm, float mag, float gx1, float gy1, float fC,
                                             float& a, float& b, int nan_compat = 0) {
    if (nan_compat && mag == 0.0f) { a = __int_as_float(0x7fc00000); b = a; return; }
    if (gm == 0.0f || mag == 0.0f) { a = 0.0f; b = 0.0f; return; }
    const float t = gm / (mag * fC);
    a = t * gx1;
    b = t * gy1;
}

}  // namespace ee
