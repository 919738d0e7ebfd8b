// ee_attack.cuh -- attack inner-loop updates and straight-through helper ops (elementwise).
//
// Replaces the sign-step / project / clamp lines of utils/attacks.py (PGD :25-27 and its ten
// copies, FGSM :121-126, CW :213-222, TRADES-L2 :391-399) and the free/fast-AT delta update
// (ImageNet/free_imagenet/AT_hfs_canny_free_imagenet_ddp.py:330-332,:314-315).  The reference
// runs 6-14 eager kernels per update; here each update is one pass: every operand is read
// once and the result written once (16 B/element for PGD L-inf).
//
// fp32 semantics are torch's: sign(0) = sign(NaN) = 0; min/max/clamp propagate NaN; the
// step alpha*sign(g) is exact, so x + alpha*sign(g) has a single rounding.
#pragma once
#include "ee_device.cuh"

namespace ee {

__device__ __forceinline__ float sgnf(float g) { return (float)((g > 0.0f) - (g < 0.0f)); }
// torch.max / torch.min (binary): NaN in either operand wins -> one FMNMX.NAN each
__device__ __forceinline__ float maxn(float a, float b) { float r; asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float minn(float a, float b) { float r; asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }

// Generic elementwise driver: NIN inputs, one output, 128-bit path when everything is 16-byte
// aligned.  Each thread keeps UNROLL independent 128-bit loads per input in flight.
template <int NIN, typename F>
__global__ void __launch_bounds__(256) ew_kernel(const float* i0, const float* i1, const float* i2, const float* i3,
                                                 const float* i4, float* out,   // out may alias an input
                                                 int64_t n, int vec_ok, F f) {
    constexpr int UNROLL = 4;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const float* in[5] = {i0, i1, i2, i3, i4};
    int64_t done = 0;
    if (vec_ok) {
        const int64_t n4 = n >> 2;
        // block-cyclic: a CTA sweeps UNROLL consecutive 256-wide float4 stripes per iteration
        for (int64_t base = (int64_t)blockIdx.x * blockDim.x * UNROLL; base < n4; base += nthreads * UNROLL) {
            float4 v[NIN][UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int64_t i = base + (int64_t)u * blockDim.x + threadIdx.x;
                if (i < n4) {
#pragma unroll
                    for (int q = 0; q < NIN; ++q) v[q][u] = __ldcs(reinterpret_cast<const float4*>(in[q]) + i);
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int64_t i = base + (int64_t)u * blockDim.x + threadIdx.x;
                if (i < n4) {
                    float a[5][4];
#pragma unroll
                    for (int q = 0; q < NIN; ++q) { a[q][0] = v[q][u].x; a[q][1] = v[q][u].y; a[q][2] = v[q][u].z; a[q][3] = v[q][u].w; }
                    float4 r;
                    r.x = f(a, 0); r.y = f(a, 1); r.z = f(a, 2); r.w = f(a, 3);
                    __stcs(reinterpret_cast<float4*>(out) + i, r);
                }
            }
        }
        done = n4 << 2;
    }
    for (int64_t i = done + tid; i < n; i += nthreads) {
        float a[5][4];
#pragma unroll
        for (int q = 0; q < NIN; ++q) a[q][0] = in[q][i];
        out[i] = f(a, 0);
    }
}

// ---- functors: a[q][k] is element k of input q ------------------------------------------------
struct FPgdLinf {     // inputs: x, g, x0
    float alpha_signed, eps, lo, hi;
    __device__ __forceinline__ float operator()(const float (&a)[5][4], int k) const {
        float t = a[0][k] + alpha_signed * sgnf(a[1][k]);
        t = minn(maxn(t, a[2][k] - eps), a[2][k] + eps);
        return minn(maxn(t, lo), hi);
    }
};
struct FFgsm {        // inputs: x, g
    float alpha_signed, lo, hi;
    __device__ __forceinline__ float operator()(const float (&a)[5][4], int k) const {
        return minn(maxn(a[0][k] + alpha_signed * sgnf(a[1][k]), lo), hi);
    }
};
struct FAddClamp {    // inputs: x, noise -- the random start x = clamp(x + U(-eps,eps), 0, 1) of utils/attacks.py:15-17
    float lo, hi;
    __device__ __forceinline__ float operator()(const float (&a)[5][4], int k) const { return minn(maxn(a[0][k] + a[1][k], lo), hi); }
};
struct FCwLinf {      // inputs: adv, g, x, min_x, max_x
    float step, magnitude;
    __device__ __forceinline__ float operator()(const float (&a)[5][4], int k) const {
        float t = a[0][k] + step * sgnf(a[1][k]);
        t = maxn(minn(t, a[2][k] + magnitude), a[2][k] - magnitude);
        t = minn(maxn(t, 0.0f), 1.0f);
        return maxn(minn(t, a[4][k]), a[3][k]);
    }
};
struct FToCompareFwd { float thr; __device__ __forceinline__ float operator()(const float (&a)[5][4], int k) const { return to_compare(a[0][k], thr); } };
struct FToCompareBwd { float thr; __device__ __forceinline__ float operator()(const float (&a)[5][4], int k) const { return (a[1][k] <= thr || a[1][k] > 1.001f) ? 0.0f : a[0][k]; } };
struct FToEqFwd { __device__ __forceinline__ float operator()(const float (&a)[5][4], int k) const { return (a[0][k] == 0.5f) ? 1.0f : 0.0f; } };
struct FToEqBwd { __device__ __forceinline__ float operator()(const float (&a)[5][4], int k) const { return (a[1][k] != 0.5f) ? 0.0f : a[0][k]; } };
struct FSafeSignFwd { __device__ __forceinline__ float operator()(const float (&a)[5][4], int k) const { const float s = sgnf(a[0][k]); return (s == 0.0f) ? -1.0f : s; } };
struct FSafeSignBwd { __device__ __forceinline__ float operator()(const float (&a)[5][4], int k) const { return (fabsf(a[1][k]) > 1.001f) ? 0.0f : a[0][k]; } };

// free / fast-AT: two outputs (delta in place, x_adv), so it has its own kernel.
static __global__ void __launch_bounds__(256) free_at_kernel(float* __restrict__ delta, const float* __restrict__ g,
                                                      const float* __restrict__ x0, float* __restrict__ x_adv,
                                                      int64_t n, int vec_ok, float alpha, float eps, float lo, float hi) {
    constexpr int UNROLL = 4;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    int64_t done = 0;
    auto upd = [&](float d, float gg) { return minn(maxn(d + alpha * sgnf(gg), -eps), eps); };
    auto adv = [&](float xx, float d) { return minn(maxn(xx + d, lo), hi); };
    if (vec_ok) {
        const int64_t n4 = n >> 2;
        for (int64_t base = (int64_t)blockIdx.x * blockDim.x * UNROLL; base < n4; base += nthreads * UNROLL) {
            float4 vd[UNROLL], vg[UNROLL], vx[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int64_t i = base + (int64_t)u * blockDim.x + threadIdx.x;
                if (i < n4) {
                    vd[u] = __ldcs(reinterpret_cast<const float4*>(delta) + i);
                    vg[u] = __ldcs(reinterpret_cast<const float4*>(g) + i);
                    if (x_adv) vx[u] = __ldcs(reinterpret_cast<const float4*>(x0) + i);
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int64_t i = base + (int64_t)u * blockDim.x + threadIdx.x;
                if (i < n4) {
                    float4 d;
                    d.x = upd(vd[u].x, vg[u].x); d.y = upd(vd[u].y, vg[u].y); d.z = upd(vd[u].z, vg[u].z); d.w = upd(vd[u].w, vg[u].w);
                    __stcs(reinterpret_cast<float4*>(delta) + i, d);
                    if (x_adv) {
                        float4 r;
                        r.x = adv(vx[u].x, d.x); r.y = adv(vx[u].y, d.y); r.z = adv(vx[u].z, d.z); r.w = adv(vx[u].w, d.w);
                        __stcs(reinterpret_cast<float4*>(x_adv) + i, r);
                    }
                }
            }
        }
        done = n4 << 2;
    }
    for (int64_t i = done + tid; i < n; i += nthreads) {
        const float d = upd(delta[i], g[i]);
        delta[i] = d;
        if (x_adv) x_adv[i] = adv(x0[i], d);
    }
}

// AVmixup vertex + mix (utils/attacks.py:469-478): vertex = clamp(inputs + (x_adv - inputs)*gamma, 0, 1) in fp32, then
// out = float(inputs*w_b + vertex*(1 - w_b)) evaluated in DOUBLE like the reference (its per-sample weight is a float64
// tensor, so torch promotes the whole mix to float64 before the final .to(torch.float)).  One pass: 12 B/element
// instead of ~9 eager kernels, three of them over float64 temporaries.
static __global__ void __launch_bounds__(256) avmixup_mix_kernel(const float* __restrict__ x_adv, const float* __restrict__ inputs,
                                                          const double* __restrict__ weight, float* __restrict__ out,
                                                          int64_t n, int64_t n_per, int vec_ok, float gamma) {
    auto mix = [&](float xa, float in, double w) {
        const float vertex = minn(maxn(in + (xa - in) * gamma, 0.0f), 1.0f);
        return (float)((double)in * w + (double)vertex * (1.0 - w));
    };
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (int64_t)gridDim.x * blockDim.x;
    if (vec_ok) {                                         // n_per % 4 == 0: a float4 never straddles two samples
        for (int64_t i = tid; i < (n >> 2); i += nthreads) {
            const double w = __ldg(weight + (i << 2) / n_per);
            const float4 a = __ldcs(reinterpret_cast<const float4*>(x_adv) + i), c = __ldcs(reinterpret_cast<const float4*>(inputs) + i);
            __stcs(reinterpret_cast<float4*>(out) + i, make_float4(mix(a.x, c.x, w), mix(a.y, c.y, w), mix(a.z, c.z, w), mix(a.w, c.w, w)));
        }
    } else {
        for (int64_t i = tid; i < n; i += nthreads) out[i] = mix(x_adv[i], inputs[i], weight[i / n_per]);
    }
}

// ---------------------------------------------------------------------------------------------
// TRADES PGD-L2 step (utils/attacks.py:391-399), one CTA of 1024 threads per sample.
// Per-sample RMS norm = sqrt(mean(v^2)) (attacks.py:360-366).  Reduction order (shared with the
// oracle): lane l accumulates elements l, l+1024, ... with fmaf; warp-shuffle tree (strides
// 16..1) inside each warp, then the same tree over the 32 warp sums.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum_1024(float v, float* sh) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) v = v + __shfl_down_sync(0xffffffffu, v, s);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();                 // protect sh from the previous use
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    float t = sh[lane];
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) t = t + __shfl_down_sync(0xffffffffu, t, s);
    return __shfl_sync(0xffffffffu, t, 0);
}

static __global__ void __launch_bounds__(1024) pgd_l2_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                      const float* __restrict__ x0, float* __restrict__ out,
                                                      int64_t n_per, float step, float eps) {
    __shared__ float sh[32];
    const int64_t off = (int64_t)blockIdx.x * n_per;
    const float* xb = x + off; const float* gb = g + off; const float* x0b = x0 + off;
    float* ob = out + off;
    float acc = 0.0f;
    for (int64_t i = threadIdx.x; i < n_per; i += 1024) { const float e = gb[i]; acc = fmaf(e, e, acc); }
    const float gn = sqrtf(block_sum_1024(acc, sh) / (float)n_per) + 1e-8f;
    acc = 0.0f;
    for (int64_t i = threadIdx.x; i < n_per; i += 1024) {
        const float xa = xb[i] + step * (gb[i] / gn);
        ob[i] = xa;                               // staged in the output buffer (same thread re-reads it)
        const float e = xa - x0b[i];
        acc = fmaf(e, e, acc);
    }
    const float dn = sqrtf(block_sum_1024(acc, sh) / (float)n_per);
    const bool cond = dn > eps;
    const float scale = eps / dn;
    for (int64_t i = threadIdx.x; i < n_per; i += 1024) {
        float d = ob[i] - x0b[i];
        if (cond) d = d * scale;
        ob[i] = minn(maxn(x0b[i] + d, 0.0f), 1.0f);
    }
}

}  // namespace ee
