// ee_gf.cuh -- the `with_gf=True` option of the *_EE models: a ZERO-padded 3 x 3 Gaussian on the edge map before the
// blend (Tiny_ImageNet/models_tinyimagenet/resnet_EE.py:185-187 and its copies in the other eight model files):
//
//     x_canny = F.conv2d(x_canny, weight_gaussian, padding=1);  x = clamp(x_hfs + w * x_canny, 0, 1)
//
// No reference YAML enables it (`gf: false` everywhere), so it is not fused into the filter kernels: the caller composes
// ee_edge_fwd_f32 -> ee_gf_blend_fwd_f32 and ee_gf_blend_bwd_f32 -> ee_edge_bwd_f32.  Canonical arithmetic of the 3 x 3
// Gaussian as everywhere else (gauss3 in ee_device.cuh): P = fma(c1, m, c0*(l + r)) for the outer rows,
// Q = fma(c2, m, c1*(l + r)) for the centre row, out = (P_up + Q) + P_dn, zero outside the image.
#pragma once
#include "ee_device.cuh"

namespace ee {

struct GfArgs {
    const float* edge;    // [B,1,H,W]
    const float* base;    // [B,C,H,W]
    const float* g_out;   // bwd: [B,C,H,W]
    float* out;           // fwd: [B,C,H,W]
    float* g_edge;        // bwd: [B,1,H,W] or null
    float* g_base;        // bwd: [B,C,H,W] or null
    int B, C, H, W;
    float c0, c1, c2, w;
};

constexpr int kGfTW = 32, kGfTH = 16;     // output tile of one CTA

__device__ __forceinline__ float gf_tap3(const float* p, int stride, float c0, float c1, float c2) {
    // p points at the centre of a 3 x 3 neighbourhood in a zero-extended shared-memory plane
    const float pu = fmaf(c1, p[-stride], c0 * (p[-stride - 1] + p[-stride + 1]));
    const float qm = fmaf(c2, p[0], c1 * (p[-1] + p[1]));
    const float pd = fmaf(c1, p[stride], c0 * (p[stride - 1] + p[stride + 1]));
    return (pu + qm) + pd;
}

// forward: out_c = clamp(base_c + w * gauss(edge), 0, 1).  Edge tile + 1-pixel halo in shared memory.
static __global__ void __launch_bounds__(256) gf_blend_fwd_kernel(const GfArgs a) {
    constexpr int SW = kGfTW + 2, SH = kGfTH + 2;
    __shared__ float se[SH * SW];
    const int H = a.H, W = a.W;
    const int b = blockIdx.z, r0 = blockIdx.y * kGfTH, c0i = blockIdx.x * kGfTW;
    const float* eb = a.edge + (size_t)b * H * W;
    for (int i = threadIdx.x; i < SH * SW; i += blockDim.x) {
        const int r = r0 - 1 + i / SW, c = c0i - 1 + i % SW;
        se[i] = (r >= 0 && r < H && c >= 0 && c < W) ? __ldg(eb + (size_t)r * W + c) : 0.0f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kGfTH * kGfTW; i += blockDim.x) {
        const int lr = i / kGfTW, lc = i % kGfTW, r = r0 + lr, c = c0i + lc;
        if (r >= H || c >= W) continue;
        const float ge = gf_tap3(se + (lr + 1) * SW + lc + 1, SW, a.c0, a.c1, a.c2);
        const float we = a.w * ge;
        for (int ch = 0; ch < a.C; ++ch) {
            const size_t o = (((size_t)b * a.C + ch) * H + r) * W + c;
            a.out[o] = clamp01_nan(__ldg(a.base + o) + we);
        }
    }
}

// backward: g_pre_c = g_out_c * [0 <= base_c + w*gauss(edge) <= 1] (= g_base_c); t = sum_c w * g_pre_c (fma chain, like
// the fused edge kernels); g_edge = gauss^T(t) = gauss(t) (symmetric taps, zero padding).  Edge tile + 2-pixel halo and t
// tile + 1-pixel halo in shared memory; g_base is written by the CTA that owns the pixel.
static __global__ void __launch_bounds__(256) gf_blend_bwd_kernel(const GfArgs a) {
    constexpr int EW = kGfTW + 4, EH = kGfTH + 4, TW = kGfTW + 2, TH = kGfTH + 2;
    __shared__ float se[EH * EW];
    __shared__ float st[TH * TW];
    const int H = a.H, W = a.W;
    const int b = blockIdx.z, r0 = blockIdx.y * kGfTH, c0i = blockIdx.x * kGfTW;
    const float* eb = a.edge + (size_t)b * H * W;
    for (int i = threadIdx.x; i < EH * EW; i += blockDim.x) {
        const int r = r0 - 2 + i / EW, c = c0i - 2 + i % EW;
        se[i] = (r >= 0 && r < H && c >= 0 && c < W) ? __ldg(eb + (size_t)r * W + c) : 0.0f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < TH * TW; i += blockDim.x) {
        const int lr = i / TW, lc = i % TW, r = r0 - 1 + lr, c = c0i - 1 + lc;
        float t = 0.0f;
        if (r >= 0 && r < H && c >= 0 && c < W) {
            const float ge = gf_tap3(se + (lr + 1) * EW + lc + 1, EW, a.c0, a.c1, a.c2);
            const float we = a.w * ge;
            const bool mine = (lr >= 1 && lr <= kGfTH && lc >= 1 && lc <= kGfTW);
            for (int ch = 0; ch < a.C; ++ch) {
                const size_t o = (((size_t)b * a.C + ch) * H + r) * W + c;
                const float pre = __ldg(a.base + o) + we;
                const float gp = (pre >= 0.0f && pre <= 1.0f) ? __ldg(a.g_out + o) : 0.0f;     // clamp backward, inclusive
                t = (ch == 0) ? gp * a.w : fmaf(gp, a.w, t);
                if (mine && a.g_base) a.g_base[o] = gp;
            }
        }
        st[i] = t;
    }
    if (!a.g_edge) return;
    __syncthreads();
    for (int i = threadIdx.x; i < kGfTH * kGfTW; i += blockDim.x) {
        const int lr = i / kGfTW, lc = i % kGfTW, r = r0 + lr, c = c0i + lc;
        if (r >= H || c >= W) continue;
        a.g_edge[((size_t)b * H + r) * W + c] = gf_tap3(st + (lr + 1) * TW + lc + 1, TW, a.c0, a.c1, a.c2);
    }
}

}  // namespace ee
