// ee_square.cuh -- Add_Square (utils/core.py:589-655), the random stripe + square perturbation the
// *_square models prepend to their input (Tiny_ImageNet/models_tinyimagenet/resnet_EE_square.py:187-189),
// forward and adjoint as ONE elementwise pass each (the reference runs ~5 + 7*n_queries eager kernels).
//
// The random numbers stay on the host side exactly as in the reference (torch CPU generator, same
// calls in the same order); the kernel receives them as two small tensors:
//   stripe[B,C,W] : sign(2*rand-1) of core.py:641, broadcast over rows
//   table[n_sq][2+C] (floats): square origin vh (row == column, core.py:649), side s, and the per-channel
//                    value 2*eps*sign (core.py:649-650) of every query
// forward  : t = clamp(x + eps*stripe, 0, 1)
//            per query: t = clamp(min(max(t + d_q, x - eps), x + eps), 0, 1)            (core.py:652-655)
// backward : g_x = g * d(out)/d(x).  x enters through t AND through both projection bounds; torch's binary
//            max / min split the gradient evenly on ties (very common here: outside the square t == x -+ eps
//            exactly), clamp passes it on [0, 1] inclusive.  The multiplier is recomputed from x, no saved tensors.
#pragma once
#include "ee_attack.cuh"

namespace ee {

struct SquareArgs {
    const float* x;        // [B,C,H,W]
    const float* stripe;   // [B,C,W]
    const float* table;    // [n_sq][2+C]
    const float* g;        // backward only: dL/d(out)
    float* out;            // forward: out ; backward: g_x
    int C, H, W, n_sq;
    int64_t n;             // B*C*H*W
    float eps;
};

template <bool BWD>
__device__ __forceinline__ float square_elem(const SquareArgs& a, float x, float sgn, float g, int c, int h, int w) {
    const float a0 = x + a.eps * sgn;
    float t = minn(maxn(a0, 0.0f), 1.0f);
    float dm = (a0 >= 0.0f && a0 <= 1.0f) ? 1.0f : 0.0f;
    const float lo = x - a.eps, hi = x + a.eps;
    const int stride = 2 + a.C;
    for (int q = 0; q < a.n_sq; ++q) {
        const float* row = a.table + (size_t)q * stride;
        const int pos = (int)__ldg(row), side = (int)__ldg(row + 1);
        const bool inside = (h >= pos) && (h < pos + side) && (w >= pos) && (w < pos + side);
        const float a2 = t + (inside ? __ldg(row + 2 + c) : 0.0f);
        if (BWD) dm = (a2 > lo) ? dm : ((a2 < lo) ? 1.0f : fmaf(0.5f, dm, 0.5f));
        const float a3 = maxn(a2, lo);
        if (BWD) dm = (a3 < hi) ? dm : ((a3 > hi) ? 1.0f : fmaf(0.5f, dm, 0.5f));
        const float a4 = minn(a3, hi);
        if (BWD) dm = (a4 >= 0.0f && a4 <= 1.0f) ? dm : 0.0f;
        t = minn(maxn(a4, 0.0f), 1.0f);
    }
    return BWD ? g * dm : t;
}

template <bool BWD, int VEC>
__global__ void __launch_bounds__(256) add_square_kernel(const SquareArgs a) {
    const int64_t nv = a.n / VEC;
    const int Wv = a.W / VEC;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t rowid = i / Wv;                 // (b*C + c)*H + h
        const int wv = (int)(i - rowid * Wv);
        const int64_t plane = rowid / a.H;            // b*C + c
        const int h = (int)(rowid - plane * a.H);
        const int c = (int)(plane % a.C);
        const float* ps = a.stripe + plane * a.W + (int64_t)wv * VEC;
        if (VEC == 4) {
            const float4 xv = __ldcs(reinterpret_cast<const float4*>(a.x) + i);
            const float4 sv = __ldg(reinterpret_cast<const float4*>(ps));
            float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (BWD) gv = __ldcs(reinterpret_cast<const float4*>(a.g) + i);
            float4 r;
            r.x = square_elem<BWD>(a, xv.x, sv.x, gv.x, c, h, wv * 4);
            r.y = square_elem<BWD>(a, xv.y, sv.y, gv.y, c, h, wv * 4 + 1);
            r.z = square_elem<BWD>(a, xv.z, sv.z, gv.z, c, h, wv * 4 + 2);
            r.w = square_elem<BWD>(a, xv.w, sv.w, gv.w, c, h, wv * 4 + 3);
            __stcs(reinterpret_cast<float4*>(a.out) + i, r);
        } else {
            a.out[i] = square_elem<BWD>(a, a.x[i], __ldg(ps), BWD ? a.g[i] : 0.0f, c, h, wv);
        }
    }
}

}  // namespace ee
