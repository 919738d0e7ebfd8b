// ee_edge_canny.cuh -- fused kernels for the full CannyFilter and CannyFilter_BPDA (+ blend).
//
// Replaces (reference paths): utils/core.py:222-326 (CannyFilter.forward), :426-505
// (CannyFilter_BPDA.forward), the blend of the *_EE models and the autograd graph through them,
// including BinaryConnectDeterministic / To_compare / To_eq backward (core.py:138-145,:350-382).
//
// Same strip decomposition as ee_edge_step125.cuh, with more planes:
//   S -> Bl -> (M = gated magnitude, META = orientation bin) -> NMS + double threshold (META gets
//   low+high / high / removed bits) -> hysteresis (3x3 sum of low+high, zero padded) -> edge.
// META is one 32-bit word per pixel:
//   bits 0-3 : NMS direction pair (orientation bin mod 4) + 1         core.py:258-260,:270
//   bits 4-5 : low + high  (2*t of core.py:315)                        core.py:300-315 / :486-492
//   bit  6   : high
//   bit  7   : removed by non-maximum suppression                      core.py:275-290 / :463-480
#pragma once
#include "ee_edge_step125.cuh"

namespace ee {

// forward: R1 = S -> M (TH+8 rows), R2 = Bl (TH+6), R3 = META (TH+4)
constexpr int kCannyFwdRowsPerTH = 3;
constexpr int kCannyFwdRowsFixed = 18;
// backward: R1 = S -> M (TH+12), R2 = Bl -> GB (TH+10), R3 = META (TH+8), R4 = A (TH+4), R5 = Bv (TH+4)
constexpr int kCannyBwdRowsPerTH = 5;
constexpr int kCannyBwdRowsFixed = 38;

enum { MODE_RAW = 0, MODE_LOW = 1, MODE_MIX = 2, MODE_HYST = 3 };

__device__ __forceinline__ int canny_mode(const EdgeArgs& a) {
    const bool bpda = (a.variant == 2);
    if (!a.has_low || (bpda && !a.has_high)) return MODE_RAW;   // core.py:326 ; BPDA has no `else: low` branch
    if (!a.has_high) return MODE_LOW;                            // core.py:323-324
    return a.hyst ? MODE_HYST : MODE_MIX;                        // core.py:315-321
}

__device__ __forceinline__ int meta_bin(int m) { return (m & 15) - 1; }
__device__ __forceinline__ int meta_lh(int m) { return (m >> 4) & 3; }
__device__ __forceinline__ int meta_hi(int m) { return (m >> 6) & 1; }
__device__ __forceinline__ int meta_removed(int m) { return (m >> 7) & 1; }

// M / META planes on rows [lo,hi): gated magnitude and orientation bin from the blurred plane
template <int VEC>
__device__ __forceinline__ void stage_mag_bin(const EdgeArgs& a, const float* Bl, int b_lo, float* M, float* META,
                                              int lo, int hi, int G, int tx, int ty) {
    const int W = a.W;
    const bool gate = (a.variant == 1);      // only CannyFilter applies alpha (core.py:263-264)
    EE_FOR_TILE(lo, hi) {
        const int col = g * VEC;
        float gx1[VEC], gy1[VEC], mm[VEC], mt[VEC];
        sobel_at<VEC>(a, Bl, b_lo, row, col, gx1, gy1);
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            const float mag = magnitude(gx1[k], gy1[k]);
            mm[k] = (gate && mag < a.alpha) ? 0.0f : mag;
            mt[k] = __int_as_float(orient_dir(gx1[k], gy1[k]) + 1);
        }
        st_vec<VEC>(M + (size_t)(row - lo) * W + col, mm);
        st_vec<VEC>(META + (size_t)(row - lo) * W + col, mt);
    }
}

// Non-maximum suppression + thresholds for VEC pixels at (row, col).  M / META planes start at
// row m_lo.  Returns the thinned magnitude and the updated META words.
template <int VEC>
__device__ __forceinline__ void nms_threshold(const EdgeArgs& a, const float* M, const float* META, int m_lo,
                                              int row, int col, float (&thin)[VEC], int (&meta)[VEC]) {
    const int W = a.W, H = a.H;
    const bool bpda = (a.variant == 2);
    float u[VEC + 2], m[VEC + 2], d[VEC + 2], mt[VEC];
    load_win<VEC, true>(zrow(M, m_lo, row - 1, H, W), col, W, u);     // directional conv is zero padded
    load_win<VEC, true>(zrow(M, m_lo, row, H, W), col, W, m);
    load_win<VEC, true>(zrow(M, m_lo, row + 1, H, W), col, W, d);
    ld_vec<VEC>(META + (size_t)(row - m_lo) * W + col, mt);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        int w = __float_as_int(mt[k]) & 15;
        const int dir = w - 1;
        const float mc = m[k + 1];
        int removed = 0;
        {
            // -1 tap offsets (row,col): 0:(0,+1) 1:(-1,+1) 2:(-1,0) 3:(-1,-1) | 4:(0,-1) 5:(+1,-1) 6:(+1,0) 7:(+1,+1)
            const float n1 = (dir == 0) ? m[k + 2] : (dir == 1) ? u[k + 2] : (dir == 2) ? u[k + 1] : u[k];
            const float n2 = (dir == 0) ? m[k] : (dir == 1) ? d[k] : (dir == 2) ? d[k + 1] : d[k + 2];
            const float d1 = mc - n1, d2 = mc - n2;
            removed = !(fminf(d1, d2) > 0.0f);
        }
        const float th = removed ? 0.0f : mc;
        thin[k] = th;
        const float lo = bpda ? to_compare(th, a.low) : sign_step(th, a.low);
        const float hi = bpda ? to_compare(th, a.high) : sign_step(th, a.high);
        const int ilo = (lo == 1.0f), ihi = (hi == 1.0f);
        meta[k] = w | ((ilo + ihi) << 4) | (ihi << 6) | (removed << 7);
    }
}

// weak_is_high for VEC pixels at (row,col): weak = (t == .5) and 1.25 * sum3x3(t) > 1  <=>
// (low+high == 1) and sum3x3(low+high) >= 2   (core.py:319-320 / :497-502; conv is zero padded)
template <int VEC>
__device__ __forceinline__ void weak_is_high(const EdgeArgs& a, const float* META, int m_lo, int row, int col,
                                             int (&wih)[VEC], int (&centre)[VEC]) {
    const int W = a.W, H = a.H;
    float u[VEC + 2], m[VEC + 2], d[VEC + 2];
    load_win<VEC, true>(zrow(META, m_lo, row - 1, H, W), col, W, u);
    load_win<VEC, true>(zrow(META, m_lo, row, H, W), col, W, m);
    load_win<VEC, true>(zrow(META, m_lo, row + 1, H, W), col, W, d);
    int cs[VEC + 2];
#pragma unroll
    for (int k = 0; k < VEC + 2; ++k)
        cs[k] = meta_lh(__float_as_int(u[k])) + meta_lh(__float_as_int(m[k])) + meta_lh(__float_as_int(d[k]));
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        const int c = __float_as_int(m[k + 1]);
        const int n = cs[k] + cs[k + 1] + cs[k + 2];
        centre[k] = c;
        wih[k] = (meta_lh(c) == 1) && (n >= 2);
    }
}

// edge value of one pixel for the non-hysteresis modes
__device__ __forceinline__ float edge_value_simple(int mode, float thin, int meta) {
    if (mode == MODE_RAW) return thin;
    if (mode == MODE_LOW) return (float)(meta_lh(meta) - meta_hi(meta));          // low only
    return (float)(meta_lh(meta) - meta_hi(meta)) * 0.5f + (float)meta_hi(meta) * 0.5f;   // low*.5 + high*.5
}

// write edge (and the blended image) for VEC pixels
template <int VEC, int NC, bool BLEND>
__device__ __forceinline__ void emit(const EdgeArgs& a, int b, int row, int col, const float (&e)[VEC]) {
    const int W = a.W;
    const int C = NC ? NC : a.C;
    (void)W;
    if (a.edge) stg_vec<VEC>(a.edge + at(a.sedge, b, 0, row, col), e);
    if (BLEND) {
        float we[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) we[k] = a.w * e[k];
        for (int c = 0; c < C; ++c) {
            float t[VEC], o[VEC];
            ldg_vec<VEC>(a.base + at(a.sbase, b, c, row, col), t);
#pragma unroll
            for (int k = 0; k < VEC; ++k) o[k] = clamp01_nan(t[k] + we[k]);
            stg_vec<VEC>(a.out + at(a.sout, b, c, row, col), o);
        }
    }
}

// -------------------------------------------------------------------------------------------
// forward
// -------------------------------------------------------------------------------------------
template <int VEC, int NC, bool BLEND>
__global__ void __launch_bounds__(256) edge_fwd_canny_kernel(const EdgeArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int b = blockIdx.x / a.tiles_per_img;
    const int ti = blockIdx.x - b * a.tiles_per_img;
    const int H = a.H, W = a.W;
    const int C = NC ? NC : a.C;
    const int r0 = ti * a.TH, r1 = min(r0 + a.TH, H);
    const int G = (W + VEC - 1) / VEC;
    const int tx = threadIdx.x % a.GX, ty = threadIdx.x / a.GX;
    const size_t hw = (size_t)H * W;
    const int mode = canny_mode(a);
    const int hc = (mode == MODE_HYST) ? 1 : 0;      // extra halo row for the hysteresis conv

    float* R1 = smem;
    float* R2 = R1 + (size_t)(a.TH + 8) * W;
    float* R3 = R2 + (size_t)(a.TH + 6) * W;

    const int s_lo = max(r0 - 3 - hc, 0), s_hi = min(r1 + 3 + hc, H);
    const int b_lo = max(r0 - 2 - hc, 0), b_hi = min(r1 + 2 + hc, H);
    const int m_lo = max(r0 - 1 - hc, 0), m_hi = min(r1 + 1 + hc, H);

    float* S = R1; float* Bl = R2;
    stage_channel_sum<VEC, NC>(a, a.x + (size_t)((int64_t)b * a.sx.b), S, s_lo, s_hi, G, tx, ty);
    __syncthreads();
    stage_blur<VEC>(a, S, s_lo, Bl, b_lo, b_hi, G, tx, ty);
    __syncthreads();
    float* M = R1; float* META = R3;
    stage_mag_bin<VEC>(a, Bl, b_lo, M, META, m_lo, m_hi, G, tx, ty);
    __syncthreads();

    if (mode != MODE_HYST) {
        EE_FOR_TILE(r0, r1) {
            const int col = g * VEC;
            float thin[VEC], e[VEC];
            int meta[VEC];
            nms_threshold<VEC>(a, M, META, m_lo, row, col, thin, meta);
#pragma unroll
            for (int k = 0; k < VEC; ++k) e[k] = edge_value_simple(mode, thin[k], meta[k]);
            emit<VEC, NC, BLEND>(a, b, row, col, e);
        }
        return;
    }
    // NMS + thresholds on rows [r0-1, r1+1): META updated in place (each thread owns its words)
    const int c_lo = max(r0 - 1, 0), c_hi = min(r1 + 1, H);
    EE_FOR_TILE(c_lo, c_hi) {
        const int col = g * VEC;
        float thin[VEC], mt[VEC];
        int meta[VEC];
        nms_threshold<VEC>(a, M, META, m_lo, row, col, thin, meta);
#pragma unroll
        for (int k = 0; k < VEC; ++k) mt[k] = __int_as_float(meta[k]);
        st_vec<VEC>(META + (size_t)(row - m_lo) * W + col, mt);
    }
    __syncthreads();
    EE_FOR_TILE(r0, r1) {
        const int col = g * VEC;
        int wih[VEC], c[VEC];
        float e[VEC];
        weak_is_high<VEC>(a, META, m_lo, row, col, wih, c);
#pragma unroll
        for (int k = 0; k < VEC; ++k) e[k] = (float)meta_hi(c[k]) + (float)wih[k];     // core.py:321 / :503
        emit<VEC, NC, BLEND>(a, b, row, col, e);
    }
}

// dL/d(thin) from dL/d(edge) for one pixel (see oracle g_thin_of and SURVEY.md A.3); `variant` is passed
// separately so that specialised kernels can make it a compile-time constant
__device__ __forceinline__ float g_thin_of_v(int variant, float low, float high, int mode, float ge, float thin, int wih) {
    if (variant == 0) return ste_sel(ge, thin, high);      // CannyFilter_step125_1: To_compare.backward on the gated magnitude
    if (mode == MODE_RAW) return ge;
    if (variant == 1) {   // CannyFilter: (sign(.)+1)/2 with the BinaryConnect STE window
        if (mode == MODE_LOW) return bcd_sel(0.5f * ge, thin, low);
        if (mode == MODE_MIX) {
            const float h = 0.5f * ge;
            return bcd_sel(0.5f * h, thin, low) + bcd_sel(0.5f * h, thin, high);
        }
        return bcd_sel(0.5f * ge, thin, high);   // hysteresis: only `high` is differentiable
    }
    // CannyFilter_BPDA: To_compare / To_eq windows
    if (mode == MODE_MIX) {
        const float h = 0.5f * ge;
        return ste_sel(h, thin, low) + ste_sel(h, thin, high);
    }
    const float gt = ge * (float)wih;
    const float g_low = 0.5f * gt;
    const float g_high = ge + 0.5f * gt;
    return ste_sel(g_low, thin, low) + ste_sel(g_high, thin, high);
}
__device__ __forceinline__ float g_thin_of(const EdgeArgs& a, int mode, float ge, float thin, int wih) {
    return g_thin_of_v(a.variant, a.low, a.high, mode, ge, thin, wih);
}

// -------------------------------------------------------------------------------------------
// backward
// -------------------------------------------------------------------------------------------
template <int VEC, int NC, bool BLEND>
__global__ void __launch_bounds__(256) edge_bwd_canny_kernel(const EdgeArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int b = blockIdx.x / a.tiles_per_img;
    const int ti = blockIdx.x - b * a.tiles_per_img;
    const int H = a.H, W = a.W;
    const int C = NC ? NC : a.C;
    const int r0 = ti * a.TH, r1 = min(r0 + a.TH, H);
    const int G = (W + VEC - 1) / VEC;
    const int tx = threadIdx.x % a.GX, ty = threadIdx.x / a.GX;
    const size_t hw = (size_t)H * W;
    const int mode = canny_mode(a);
    const bool bpda = (a.variant == 2);
    // the forward edge value (blend) or BPDA's To_eq path need weak_is_high on the A/Bv rows
    const int hc = (mode == MODE_HYST && (BLEND || bpda)) ? 1 : 0;
    const bool want_gx = (a.g_x != nullptr);
    const int ha = want_gx ? 2 : 0;                  // halo of the A/Bv rows

    float* R1 = smem;
    float* R2 = R1 + (size_t)(a.TH + 12) * W;
    float* R3 = R2 + (size_t)(a.TH + 10) * W;
    float* R4 = R3 + (size_t)(a.TH + 8) * W;
    float* R5 = R4 + (size_t)(a.TH + 4) * W;

    const int ab_lo = max(r0 - ha, 0), ab_hi = min(r1 + ha, H);
    const int c_lo = max(r0 - ha - hc, 0), c_hi = min(r1 + ha + hc, H);          // thresholded META rows
    const int m_lo = max(r0 - ha - hc - 1, 0), m_hi = min(r1 + ha + hc + 1, H);
    const int b_lo = max(r0 - ha - hc - 2, 0), b_hi = min(r1 + ha + hc + 2, H);
    const int s_lo = max(r0 - ha - hc - 3, 0), s_hi = min(r1 + ha + hc + 3, H);
    const int gb_lo = max(r0 - 1, 0), gb_hi = min(r1 + 1, H);

    float* S = R1; float* Bl = R2;
    stage_channel_sum<VEC, NC>(a, a.x + (size_t)((int64_t)b * a.sx.b), S, s_lo, s_hi, G, tx, ty);
    __syncthreads();
    stage_blur<VEC>(a, S, s_lo, Bl, b_lo, b_hi, G, tx, ty);
    __syncthreads();
    float* M = R1; float* META = R3;
    stage_mag_bin<VEC>(a, Bl, b_lo, M, META, m_lo, m_hi, G, tx, ty);
    __syncthreads();
    EE_FOR_TILE(c_lo, c_hi) {
        const int col = g * VEC;
        float thin[VEC], mt[VEC];
        int meta[VEC];
        nms_threshold<VEC>(a, M, META, m_lo, row, col, thin, meta);
#pragma unroll
        for (int k = 0; k < VEC; ++k) mt[k] = __int_as_float(meta[k]);
        st_vec<VEC>(META + (size_t)(row - m_lo) * W + col, mt);
    }
    __syncthreads();

    float* A = R4; float* Bv = R5;
    EE_FOR_TILE(ab_lo, ab_hi) {
        const int col = g * VEC;
        const size_t pix = (size_t)row * W + col;
        float gx1[VEC], gy1[VEC], mv[VEC], thin[VEC], ge[VEC];
        int wih[VEC], meta[VEC];
        sobel_at<VEC>(a, Bl, b_lo, row, col, gx1, gy1);
        ld_vec<VEC>(M + (size_t)(row - m_lo) * W + col, mv);
        if (hc) {
            weak_is_high<VEC>(a, META, m_lo, row, col, wih, meta);
        } else {
            float mt[VEC];
            ld_vec<VEC>(META + (size_t)(row - m_lo) * W + col, mt);
#pragma unroll
            for (int k = 0; k < VEC; ++k) { meta[k] = __float_as_int(mt[k]); wih[k] = 0; }
        }
#pragma unroll
        for (int k = 0; k < VEC; ++k) thin[k] = meta_removed(meta[k]) ? 0.0f : mv[k];
        if (BLEND) {
            float we[VEC];
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                const float e = (mode == MODE_HYST) ? (float)meta_hi(meta[k]) + (float)wih[k]
                                                    : edge_value_simple(mode, thin[k], meta[k]);
                we[k] = a.w * e;
            }
            const bool interior = (row >= r0 && row < r1);
            for (int c = 0; c < C; ++c) {
                float bs[VEC], go[VEC], gp[VEC];
                ldg_vec<VEC>(a.base + at(a.sbase, b, c, row, col), bs);
                ldg_vec<VEC>(a.g_in + at(a.sg, b, c, row, col), go);
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    const float pre = bs[k] + we[k];
                    gp[k] = (pre >= 0.0f && pre <= 1.0f) ? go[k] : 0.0f;
                    ge[k] = (c == 0) ? gp[k] * a.w : fmaf(gp[k], a.w, ge[k]);
                }
                if (a.g_base && interior) stg_vec<VEC>(a.g_base + at(a.sgbase, b, c, row, col), gp);
            }
        } else {
            ldg_vec<VEC>(a.g_in + at(a.sg, b, 0, row, col), ge);
        }
        if (want_gx) {
            float av[VEC], bv[VEC];
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                const float mag = magnitude(gx1[k], gy1[k]);
                float gm = g_thin_of(a, mode, ge[k], thin[k], wih[k]);
                if (meta_removed(meta[k])) gm = 0.0f;                     // core.py:290 / :480
                if (a.variant == 1 && mag < a.alpha) gm = 0.0f;           // torch.where backward
                mag_backward(gm, mag, gx1[k], gy1[k], a.fC, av[k], bv[k], a.nan_compat);
            }
            st_vec<VEC>(A + (size_t)(row - ab_lo) * W + col, av);
            st_vec<VEC>(Bv + (size_t)(row - ab_lo) * W + col, bv);
        }
    }
    if (!want_gx) return;
    __syncthreads();
    float* GB = R2;
    stage_sobel_adjoint<VEC>(a, A, Bv, ab_lo, GB, gb_lo, gb_hi, G, tx, ty);
    __syncthreads();
    stage_gauss_adjoint_store<VEC, NC>(a, GB, gb_lo, b, r0, r1, G, tx, ty);
}

}  // namespace ee
