// ee_edge_canny_fast.cuh -- tuned kernels for the full CannyFilter and CannyFilter_BPDA (+ blend),
// built from the same pieces as ee_edge_fast.cuh (padded planes, sliding row windows, exact /3,
// width specialisation).  Same canonical arithmetic as the generic kernels of ee_edge_canny.cuh,
// which stay as the any-shape fallback and as the second implementation the tests compare with.
//
// Replaces (reference): utils/core.py:222-326 (CannyFilter), :426-505 (CannyFilter_BPDA), the blend
// of the *_EE models and the autograd through them.
//
// Planes (row stride Wp = W + 8, one pad column each side):
//   S, Bl      : replicate padded (as in the step125 kernels)
//   M          : gated gradient magnitude, ZERO padded (the directional conv is zero padded, core.py:268)
//   META       : one word per pixel, zero padded: bits 0-3 direction pair + 1, 4-5 low+high, 6 high, 7 removed
#pragma once
#ifndef EE_L2_PREFETCH_BWD_CANNY
#define EE_L2_PREFETCH_BWD_CANNY 2
#endif
#include "ee_edge_canny.cuh"
#include "ee_edge_fast.cuh"

namespace ee {

// forward : R1 = S -> M (TH+8 rows), R2 = Bl (TH+6), R3 = META (TH+4)
constexpr int kCannyFastFwdRowsPerTH = 3, kCannyFastFwdRowsFixed = 18;
// backward: R1 = S -> M -> A (TH+12), R2 = Bl -> GB (TH+10), R3 = META (TH+8), R4 = gx1 -> Bv (TH+4), R5 = gy1 (TH+4)
constexpr int kCannyFastBwdRowsPerTH = 5, kCannyFastBwdRowsFixed = 38;

__device__ __forceinline__ void win_to_array(const Win& w, float (&e)[6]) {
    e[0] = w.l; e[1] = w.m0; e[2] = w.m1; e[3] = w.m2; e[4] = w.m3; e[5] = w.r;
}
__device__ __forceinline__ Win zero_win() { Win w; w.l = w.m0 = w.m1 = w.m2 = w.m3 = w.r = 0.0f; return w; }

// ---- stage MB: M (gated magnitude) and META (direction) rows [lo,hi) from the blurred plane ----
// STORE_G (backward): also keep gx1 / gy1 of rows [g_lo, g_hi) in two planes (origin g_lo), so that the A/Bv stage
// does not recompute the Sobel pair, the division by C and the square root.
// DEFER_GY (EVEN only: one chunk per thread): gy1 is returned in registers instead; the caller stores it into the
// blurred plane's region after a barrier, so that the backward needs four planes instead of five.
template <int DIVM, int R, bool EVEN = false, bool STORE_G = false, bool DEFER_GY = false>
__device__ __forceinline__ void cfast_stage_mag_dir(const FastArgs& a, const Geo geo, const float* Bl, int b_lo, float* M,
                                                    float* META, int lo, int hi, int tx, int ty, const int variant,
                                                    float* GXp = nullptr, float* GYp = nullptr, int g_lo = 0, int g_hi = 0,
                                                    float4* gy_keep = nullptr) {
    const int W = geo.W, H = geo.H, Wp = geo.Wp;
    const float fC = a.e.fC;
    const bool gate = (variant == 1);                // only CannyFilter applies alpha (core.py:263-264)
    EE_FOR_CHUNKS_E(lo, hi) {
        const int lc = g * 4, col = geo.cs + lc, ra = lo + ch * R, rb = EVEN ? ra + R : min(ra + R, hi);
        const float* pbl = Bl + kPadL + lc;
        float D[3][4], V[3][4];
#pragma unroll
        for (int i = 0; i < R + 2; ++i) {
            const int rin = ra - 1 + i;
            if (EVEN || rin <= rb) {
                const int rc = min(max(rin, 0), H - 1);
                sobel_partials(ld_win(pbl + (rc - b_lo) * Wp), D[i % 3], V[i % 3]);
            }
            if (i >= 2 && (EVEN || ra + i - 2 < rb)) {
                float sgx[4], sgy[4], gx1[4], gy1[4], mm[4], mt[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    sgx[k] = fmaf(0.5f, D[(i - 2) % 3][k] + D[i % 3][k], D[(i - 1) % 3][k]);
                    sgy[k] = V[i % 3][k] - V[(i - 2) % 3][k];
                }
                div_channels8<DIVM>(sgx, sgy, fC, gx1, gy1);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float mag = magnitude(gx1[k], gy1[k]);
                    mm[k] = (gate && mag < a.e.alpha) ? 0.0f : mag;
                    mt[k] = __int_as_float(orient_dir(gx1[k], gy1[k]) + 1);
                }
                const int q = (ra + i - 2 - lo) * Wp + kPadL + lc;
                st_plane(M + q, mm, col == 0, col + 4 == W, 0.0f, 0.0f);
                st_plane(META + q, mt, col == 0, col + 4 == W, 0.0f, 0.0f);
                if (STORE_G) {
                    const int rg = ra + i - 2;
                    if (EVEN || (rg >= g_lo && rg < g_hi)) {
                        const int qg = (rg - g_lo) * Wp + kPadL + lc;
                        *reinterpret_cast<float4*>(GXp + qg) = make_float4(gx1[0], gx1[1], gx1[2], gx1[3]);
                        if (DEFER_GY) gy_keep[i - 2] = make_float4(gy1[0], gy1[1], gy1[2], gy1[3]);
                        else *reinterpret_cast<float4*>(GYp + qg) = make_float4(gy1[0], gy1[1], gy1[2], gy1[3]);
                    }
                }
            }
        }
    }
}

// NMS + double threshold of 4 pixels from the three M windows (rows above / centre / below) and the
// centre META words.  Returns thin[] and the updated META words.
__device__ __forceinline__ void nms_threshold4(const FastArgs& a, const Win& wu, const Win& wm, const Win& wd,
                                               const float4 mt, float (&thin)[4], int (&meta)[4], const int variant) {
    const bool bpda = (variant == 2);
    float u[6], m[6], d[6];
    win_to_array(wu, u); win_to_array(wm, m); win_to_array(wd, d);
    const int words[4] = {__float_as_int(mt.x), __float_as_int(mt.y), __float_as_int(mt.z), __float_as_int(mt.w)};
    const bool neg_lo = (0.0f > a.e.low), neg_hi = (0.0f > a.e.high);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int w = words[k] & 15;
        const int dir = w - 1;
        const float mc = m[k + 1];
        // -1 tap offsets (row,col): 0:(0,+1) 1:(-1,+1) 2:(-1,0) 3:(-1,-1) | opposite: (0,-1) (+1,-1) (+1,0) (+1,+1)
        // min(mc - n1, mc - n2) > 0  <=>  mc > max(n1, n2) (exact for finite values: a - b > 0 <=> a > b with
        // gradual underflow): the larger neighbour of each of the four pairs, then one select by direction
        const float mx0 = fmaxf(m[k + 2], m[k]), mx1 = fmaxf(u[k + 2], d[k]);
        const float mx2 = fmaxf(u[k + 1], d[k + 1]), mx3 = fmaxf(u[k], d[k + 2]);
        const float mx = (dir & 2) ? ((dir & 1) ? mx3 : mx2) : ((dir & 1) ? mx1 : mx0);
        const int removed = !(mc > mx);
        const float th = removed ? 0.0f : mc;
        thin[k] = th;
        // sign_step: (th - thr > 0) <=> th > thr ; to_compare additionally turns "not above" into 1 for a negative threshold
        const bool plo = th > a.e.low, phi = th > a.e.high;
        const int ilo = bpda ? (plo || (neg_lo && th <= a.e.low)) : plo;
        const int ihi = bpda ? (phi || (neg_hi && th <= a.e.high)) : phi;
        meta[k] = w | ((ilo + ihi) << 4) | (ihi << 6) | (removed << 7);
    }
}

// horizontal 3-sums of (low+high) of one META row: hs[k] = lh(k-1) + lh(k) + lh(k+1); also the centre words
__device__ __forceinline__ void meta_partials(const Win& w, int (&hs)[4], int (&cw)[4]) {
    const int b0 = __float_as_int(w.l), b1 = __float_as_int(w.m0), b2 = __float_as_int(w.m1), b3 = __float_as_int(w.m2),
              b4 = __float_as_int(w.m3), b5 = __float_as_int(w.r);
    const int l0 = meta_lh(b0), l1 = meta_lh(b1), l2 = meta_lh(b2), l3 = meta_lh(b3), l4 = meta_lh(b4), l5 = meta_lh(b5);
    hs[0] = l0 + l1 + l2; hs[1] = l1 + l2 + l3; hs[2] = l2 + l3 + l4; hs[3] = l3 + l4 + l5;
    cw[0] = b1; cw[1] = b2; cw[2] = b3; cw[3] = b4;
}

// write edge (+ blended image) of 4 pixels at (row, col)
template <int NC, bool BLEND, bool NHWC = false>
__device__ __forceinline__ void cfast_emit(const FastArgs& a, const Geo geo, int b, int row, int col, const float (&e)[4]) {
    const int W = geo.W;
    const int C = NC ? NC : a.e.C;
    const size_t hw = (size_t)geo.H * W;
    const int pix = row * W + col;
    if (a.e.edge) __stcs(reinterpret_cast<float4*>(a.e.edge + (size_t)b * hw + pix), make_float4(e[0], e[1], e[2], e[3]));
    if (BLEND) {
        const float w0 = a.e.w * e[0], w1 = a.e.w * e[1], w2 = a.e.w * e[2], w3 = a.e.w * e[3];
        const float* base_b = a.e.base + (size_t)b * C * hw;
        float* out_b = a.e.out + (size_t)b * C * hw;
        if (NC) {
            float4 bs[NC ? NC : 1];
            ld_px4<NC, NHWC>(base_b, hw, pix, bs);
#pragma unroll
            for (int c = 0; c < (NC ? NC : 1); ++c)
                bs[c] = make_float4(clamp01_fast(bs[c].x + w0), clamp01_fast(bs[c].y + w1), clamp01_fast(bs[c].z + w2),
                                    clamp01_fast(bs[c].w + w3));
            st_px4<NC, NHWC>(out_b, hw, pix, bs);
        } else {
            for (int c = 0; c < C; ++c) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(base_b + c * hw + pix));
                __stcs(reinterpret_cast<float4*>(out_b + c * hw + pix),
                       make_float4(clamp01_fast(t.x + w0), clamp01_fast(t.y + w1), clamp01_fast(t.z + w2), clamp01_fast(t.w + w3)));
            }
        }
    }
}

// ---- stage NMS: rows [lo,hi).  EMIT = false: update META in place (hysteresis needs a 3x3 sum of it);
//      EMIT = true: the non-hysteresis modes write their output directly. ------------------------------
template <int NC, bool BLEND, int R, bool EMIT, bool NHWC = false, bool EVEN = false>
__device__ __forceinline__ void cfast_stage_nms(const FastArgs& a, const Geo geo, const float* M, float* META, int m_lo,
                                                int lo, int hi, int b, int mode, int tx, int ty, const int variant) {
    const int H = geo.H, Wp = geo.Wp;
    EE_FOR_CHUNKS_E(lo, hi) {
        const int lc = g * 4, col = geo.cs + lc, ra = lo + ch * R, rb = EVEN ? ra + R : min(ra + R, hi);
        if (!EVEN && EMIT && (col < geo.c0 || col >= geo.c1)) continue;   // halo groups produce no output
        const float* pm = M + kPadL + lc;
        Win wm[3];
#pragma unroll
        for (int i = 0; i < R + 2; ++i) {
            const int rin = ra - 1 + i;
            if (EVEN || rin <= rb) {
                if (EVEN && i > 0 && i < R + 1) wm[i % 3] = ld_win(pm + (rin - m_lo) * Wp);
                else wm[i % 3] = (rin >= 0 && rin < H) ? ld_win(pm + (rin - m_lo) * Wp) : zero_win();
            }
            if (i >= 2 && (EVEN || ra + i - 2 < rb)) {
                const int p = ra + i - 2;
                float* pmeta = META + (p - m_lo) * Wp + kPadL + lc;
                const float4 mt = *reinterpret_cast<const float4*>(pmeta);
                float thin[4];
                int meta[4];
                nms_threshold4(a, wm[(i - 2) % 3], wm[(i - 1) % 3], wm[i % 3], mt, thin, meta, variant);
                if (EMIT) {
                    float e[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) e[k] = edge_value_simple(mode, thin[k], meta[k]);
                    cfast_emit<NC, BLEND, NHWC>(a, geo, b, p, col, e);
                } else {
                    *reinterpret_cast<float4*>(pmeta) = make_float4(__int_as_float(meta[0]), __int_as_float(meta[1]),
                                                                    __int_as_float(meta[2]), __int_as_float(meta[3]));
                }
            }
        }
    }
}

// -------------------------------------------------------------------------------------------
// forward
// -------------------------------------------------------------------------------------------
// VAR / MODE: 0 / -1 = read variant and output mode from the arguments; the hot configuration (models always
// call the filter with both thresholds and hysteresis=True) is compiled with them as constants.
#ifndef EE_MINB_CANNY_FWD
#define EE_MINB_CANNY_FWD 3
#endif
// L2 bulk prefetch of `base` (consumed by the last stage): 0 off, 1 after the x loads, 2 after the blur, 3 after mag/dir
#ifndef EE_L2_PREFETCH_FWD_CANNY
#define EE_L2_PREFETCH_FWD_CANNY 1
#endif
template <int NC, bool BLEND, int R, int WT, int WG, bool NHWC = false, int HT = 0, int VAR = 0, int MODE = -1>
__global__ void __launch_bounds__(256, EE_MINB_CANNY_FWD) edge_fwd_canny_fast(const FastArgs a) {
    extern __shared__ __align__(16) float smem[];
    constexpr int DIVM = (NC == 1) ? 0 : (NC == 3 ? 1 : 2);
    constexpr bool EVEN = (HT != 0);
    static_assert(!EVEN || (WT != 0 && WT == WG && HT % R == 0), "HT needs a single constant-width column tile");
    const int b = EVEN ? blockIdx.x : blockIdx.x / a.e.tiles_per_img;
    const int tq = EVEN ? 0 : blockIdx.x - b * a.e.tiles_per_img;
    const int ti = EVEN ? 0 : tq / a.tiles_x;                        // row-strip index; column tile = tq % tiles_x
    Geo geo = make_geo<WT, WG>(a, tq - ti * a.tiles_x);
    if (EVEN) { geo.H = HT; geo.RY = HT / R; }
    const int H = geo.H, W = geo.W, Wp = geo.Wp;
    const int C = NC ? NC : a.e.C;
    const int r0 = EVEN ? 0 : ti * a.e.TH, r1 = EVEN ? HT : min(r0 + a.e.TH, H);
    const int tx = threadIdx.x % geo.GX, ty = threadIdx.x / geo.GX;
    const size_t hw = (size_t)H * W;
    const bool active = ty < geo.RY;
    const int variant = VAR ? VAR : a.e.variant;
    const int mode = (MODE >= 0) ? MODE : canny_mode(a.e);
    const int hc = (mode == MODE_HYST) ? 1 : 0;
    const bool one_tile = EVEN || a.tiles_x == 1;

    float* R1 = smem;
    float* R2 = R1 + (size_t)(EVEN ? HT : min(a.e.TH + 8, H)) * Wp;
    float* R3 = R2 + (size_t)(EVEN ? HT : min(a.e.TH + 6, H)) * Wp;
    const int s_lo = max(r0 - 3 - hc, 0), s_hi = min(r1 + 3 + hc, H);
    const int b_lo = max(r0 - 2 - hc, 0), b_hi = min(r1 + 2 + hc, H);
    const int m_lo = max(r0 - 1 - hc, 0), m_hi = min(r1 + 1 + hc, H);

    float* S = R1; float* Bl = R2;
    auto prefetch_base = [&]() {
        if (BLEND && (!NHWC || (r0 == 0 && r1 == H)) && one_tile && threadIdx.x < 32 && C <= 32) prefetch_rows(a.e.base, b, C, H, W, r0, r1, threadIdx.x);
    };
    if (active) fast_stage_sum<NC, R, NHWC, EVEN>(a, geo, a.e.x + (size_t)b * C * hw, S, s_lo, s_hi, tx, ty);
#if EE_L2_PREFETCH_FWD_CANNY == 1
    prefetch_base();
#endif
    __syncthreads();
    if (active) fast_stage_blur<R, EVEN>(a, geo, S, s_lo, Bl, b_lo, b_hi, tx, ty);
#if EE_L2_PREFETCH_FWD_CANNY == 2
    prefetch_base();
#endif
    __syncthreads();
    float* M = R1; float* META = R3;
    if (active) cfast_stage_mag_dir<DIVM, R, EVEN>(a, geo, Bl, b_lo, M, META, m_lo, m_hi, tx, ty, variant);
#if EE_L2_PREFETCH_FWD_CANNY == 3
    prefetch_base();
#endif
    __syncthreads();
    if (mode != MODE_HYST) {
        if (active) cfast_stage_nms<NC, BLEND, R, true, NHWC, EVEN>(a, geo, M, META, m_lo, r0, r1, b, mode, tx, ty, variant);
        return;
    }
    const int c_lo = max(r0 - 1, 0), c_hi = min(r1 + 1, H);
    if (active) cfast_stage_nms<NC, BLEND, R, false, NHWC, EVEN>(a, geo, M, META, m_lo, c_lo, c_hi, b, mode, tx, ty, variant);
    __syncthreads();
    if (!active) return;
    // hysteresis (core.py:317-321 / :494-503): weak = (low+high == 1), kept if the zero-padded 3x3 sum of
    // (low+high) is >= 2; edge = high + weak_is_high
    EE_FOR_CHUNKS_E(r0, r1) {
        const int lc = g * 4, col = geo.cs + lc, ra = r0 + ch * R, rb = EVEN ? ra + R : min(ra + R, r1);
        if (!EVEN && (col < geo.c0 || col >= geo.c1)) continue;  // halo groups produce no output
        const float* pmt = META + kPadL + lc;
        int hs[3][4], cw[3][4];
#pragma unroll
        for (int i = 0; i < R + 2; ++i) {
            const int rin = ra - 1 + i;
            if (EVEN || rin <= rb) {
                if (EVEN && i > 0 && i < R + 1) meta_partials(ld_win(pmt + (rin - m_lo) * Wp), hs[i % 3], cw[i % 3]);
                else meta_partials((rin >= 0 && rin < H) ? ld_win(pmt + (rin - m_lo) * Wp) : zero_win(), hs[i % 3], cw[i % 3]);
            }
            if (i >= 2 && (EVEN || ra + i - 2 < rb)) {
                float e[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int n = hs[(i - 2) % 3][k] + hs[(i - 1) % 3][k] + hs[i % 3][k];
                    const int c = cw[(i - 1) % 3][k];
                    const int wih = (meta_lh(c) == 1) && (n >= 2);
                    e[k] = (float)(meta_hi(c) + wih);
                }
                cfast_emit<NC, BLEND, NHWC>(a, geo, b, ra + i - 2, col, e);
            }
        }
    }
}

// -------------------------------------------------------------------------------------------
// backward
// -------------------------------------------------------------------------------------------
template <int NC, bool BLEND, int R, int WT, int WG, bool NHWC = false, int HT = 0, int VAR = 0, int MODE = -1>
#ifndef EE_MINB_CANNY_BWD
#define EE_MINB_CANNY_BWD 2
#endif
#ifndef EE_MINB_CANNY_BWD_EVEN
#define EE_MINB_CANNY_BWD_EVEN 3
#endif
__global__ void __launch_bounds__(256, HT ? EE_MINB_CANNY_BWD_EVEN : EE_MINB_CANNY_BWD) edge_bwd_canny_fast(const FastArgs a) {
    extern __shared__ __align__(16) float smem[];
    constexpr int DIVM = (NC == 1) ? 0 : (NC == 3 ? 1 : 2);
    constexpr bool EVEN = (HT != 0);
    static_assert(!EVEN || (WT != 0 && WT == WG && HT % R == 0), "HT needs a single constant-width column tile");
    const int b = EVEN ? blockIdx.x : blockIdx.x / a.e.tiles_per_img;
    const int tq = EVEN ? 0 : blockIdx.x - b * a.e.tiles_per_img;
    const int ti = EVEN ? 0 : tq / a.tiles_x;                        // row-strip index; column tile = tq % tiles_x
    Geo geo = make_geo<WT, WG>(a, tq - ti * a.tiles_x);
    if (EVEN) { geo.H = HT; geo.RY = HT / R; }
    const int H = geo.H, W = geo.W, Wp = geo.Wp;
    const int C = NC ? NC : a.e.C;
    const int r0 = EVEN ? 0 : ti * a.e.TH, r1 = EVEN ? HT : min(r0 + a.e.TH, H);
    const int tx = threadIdx.x % geo.GX, ty = threadIdx.x / geo.GX;
    const size_t hw = (size_t)H * W;
    const bool active = ty < geo.RY;
    const int variant = VAR ? VAR : a.e.variant;
    const int mode = (MODE >= 0) ? MODE : canny_mode(a.e);
    const bool bpda = (variant == 2);
    // the forward edge value (blend) or BPDA's To_eq path need weak_is_high on the A/Bv rows
    const int hc = (mode == MODE_HYST && (BLEND || bpda)) ? 1 : 0;
    const bool want_gx = (a.e.g_x != nullptr);
    const int ha = want_gx ? 2 : 0;
    const float fC = a.e.fC, wgt = a.e.w;

    float* R1 = smem;
    float* R2 = R1 + (size_t)(EVEN ? HT : min(a.e.TH + 12, H)) * Wp;
    float* R3 = R2 + (size_t)(EVEN ? HT : min(a.e.TH + 10, H)) * Wp;
    float* R4 = R3 + (size_t)(EVEN ? HT : min(a.e.TH + 8, H)) * Wp;
    float* R5 = R4 + (size_t)(EVEN ? HT : min(a.e.TH + 4, H)) * Wp;

    const int ab_lo = max(r0 - ha, 0), ab_hi = min(r1 + ha, H);
    const int c_lo = max(r0 - ha - hc, 0), c_hi = min(r1 + ha + hc, H);
    const int m_lo = max(r0 - ha - hc - 1, 0), m_hi = min(r1 + ha + hc + 1, H);
    const int b_lo = max(r0 - ha - hc - 2, 0), b_hi = min(r1 + ha + hc + 2, H);
    const int s_lo = max(r0 - ha - hc - 3, 0), s_hi = min(r1 + ha + hc + 3, H);
    const int gb_lo = max(r0 - 1, 0), gb_hi = min(r1 + 1, H);

    float* S = R1; float* Bl = R2;
    if (active) fast_stage_sum<NC, R, NHWC, EVEN>(a, geo, a.e.x + (size_t)b * C * hw, S, s_lo, s_hi, tx, ty);
    // L2 bulk prefetch of the A/Bv stage's operands (base, g_out).  Placement (EE_L2_PREFETCH_BWD_CANNY): 0 off,
    // 1 right after the x loads (four stages ahead: measured -8 % at 2 CTAs/SM), 2 after the blur, 3 after mag/dir
    auto prefetch_ab_operands = [&]() {
        if (C <= 32 && (EVEN || a.tiles_x == 1) && (!NHWC || (ab_lo == 0 && ab_hi == H))) {
            if (BLEND) {
                if (threadIdx.x < 32) prefetch_rows(a.e.base, b, C, H, W, ab_lo, ab_hi, threadIdx.x);
                else if (threadIdx.x < 64) prefetch_rows(a.e.g_in, b, C, H, W, ab_lo, ab_hi, threadIdx.x - 32);
            } else if (threadIdx.x == 0) {
                prefetch_rows(a.e.g_in, b, 1, H, W, ab_lo, ab_hi, 0);
            }
        }
    };
#if EE_L2_PREFETCH_BWD_CANNY == 1
    prefetch_ab_operands();
#endif
    __syncthreads();
    if (active) fast_stage_blur<R, EVEN>(a, geo, S, s_lo, Bl, b_lo, b_hi, tx, ty);
#if EE_L2_PREFETCH_BWD_CANNY == 2
    prefetch_ab_operands();
#endif
    __syncthreads();
    float* M = R1; float* META = R3;
    // gy1 plane: whole-image tiles keep it in registers over a barrier and then reuse the (dead) blurred plane's
    // region, so they need 4 planes (3 CTAs/SM at 64x64); strips use a fifth region
    float* GXp = R4; float* GYp = EVEN ? R2 : R5;
    if (EVEN) {
        float4 gy_keep[R];
        if (active) cfast_stage_mag_dir<DIVM, R, EVEN, true, true>(a, geo, Bl, b_lo, M, META, m_lo, m_hi, tx, ty, variant, GXp, GYp, ab_lo, ab_hi, gy_keep);
        __syncthreads();                        // every thread has read its Bl windows
        if (active) {
#pragma unroll
            for (int i = 0; i < R; ++i) *reinterpret_cast<float4*>(GYp + (ty * R + i) * Wp + kPadL + tx * 4) = gy_keep[i];
        }
    } else if (active) {
        cfast_stage_mag_dir<DIVM, R, EVEN, true>(a, geo, Bl, b_lo, M, META, m_lo, m_hi, tx, ty, variant, GXp, GYp, ab_lo, ab_hi);
    }
#if EE_L2_PREFETCH_BWD_CANNY == 3
    prefetch_ab_operands();
#endif
    __syncthreads();
    if (active) cfast_stage_nms<NC, false, R, false, false, EVEN>(a, geo, M, META, m_lo, c_lo, c_hi, b, mode, tx, ty, variant);
    __syncthreads();

    // ---- A / Bv on rows [ab_lo, ab_hi).  gx1 / gy1 come from the planes the mag/dir stage kept, the gated magnitude
    //      from M (M == 0 where the alpha gate closed or mag == 0: both give a zero gradient).  A overwrites M and Bv
    //      overwrites gx1 IN PLACE (same thread, same element; M and gx1 are only read at the thread's own pixels in
    //      this stage), so A's origin is M's row of ab_lo. -------------------------------------------------------
    float* A = R1 + (size_t)(ab_lo - m_lo) * Wp; float* Bv = R4;
    if (active) {
        const float* base_b = a.e.base + (size_t)b * C * hw;
        const float* gin_b = a.e.g_in + (size_t)b * (BLEND ? C : 1) * hw;
        float* gbase_b = a.e.g_base ? a.e.g_base + (size_t)b * C * hw : nullptr;
        EE_FOR_CHUNKS_E(ab_lo, ab_hi) {
            const int lc = g * 4, col = geo.cs + lc, ra = ab_lo + ch * R, rb = EVEN ? ra + R : min(ra + R, ab_hi);
            const float* pmt = META + kPadL + lc;
            int hs[3][4], cw[3][4];
#pragma unroll
            for (int i = 0; i < R + 2; ++i) {
                const int rin = ra - 1 + i;
                if (EVEN || rin <= rb) {
                    if (hc) {
                        if (EVEN && i > 0 && i < R + 1) meta_partials(ld_win(pmt + (rin - m_lo) * Wp), hs[i % 3], cw[i % 3]);
                        else meta_partials((rin >= 0 && rin < H) ? ld_win(pmt + (rin - m_lo) * Wp) : zero_win(), hs[i % 3], cw[i % 3]);
                    } else if (EVEN ? (i > 0 && i < R + 1) : (rin >= ra && rin < rb)) {
                        const float4 t = *reinterpret_cast<const float4*>(pmt + (rin - m_lo) * Wp);
                        cw[i % 3][0] = __float_as_int(t.x); cw[i % 3][1] = __float_as_int(t.y);
                        cw[i % 3][2] = __float_as_int(t.z); cw[i % 3][3] = __float_as_int(t.w);
                    }
                }
                if (i >= 2 && (EVEN || ra + i - 2 < rb)) {
                    const int rout = ra + i - 2;
                    const int pix = rout * W + col;
                    const int q = (rout - ab_lo) * Wp + kPadL + lc;
                    float gx1[4], gy1[4], mag[4], thin[4], ge[4];
                    int meta[4], wih[4];
                    {
                        const float4 tx4 = *reinterpret_cast<const float4*>(GXp + q);
                        const float4 ty4 = *reinterpret_cast<const float4*>(GYp + q);
                        const float4 tm4 = *reinterpret_cast<const float4*>(A + q);          // = M of this row
                        gx1[0] = tx4.x; gx1[1] = tx4.y; gx1[2] = tx4.z; gx1[3] = tx4.w;
                        gy1[0] = ty4.x; gy1[1] = ty4.y; gy1[2] = ty4.z; gy1[3] = ty4.w;
                        mag[0] = tm4.x; mag[1] = tm4.y; mag[2] = tm4.z; mag[3] = tm4.w;
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        meta[k] = cw[(i - 1) % 3][k];
                        thin[k] = meta_removed(meta[k]) ? 0.0f : mag[k];
                        wih[k] = 0;
                        if (hc) {
                            const int n = hs[(i - 2) % 3][k] + hs[(i - 1) % 3][k] + hs[i % 3][k];
                            wih[k] = (meta_lh(meta[k]) == 1) && (n >= 2);
                        }
                    }
                    if (BLEND) {
                        float we[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float e = (mode == MODE_HYST) ? (float)(meta_hi(meta[k]) + wih[k])
                                                                : edge_value_simple(mode, thin[k], meta[k]);
                            we[k] = wgt * e;
                        }
                        const bool interior = EVEN || (rout >= r0 && rout < r1 && col >= geo.c0 && col < geo.c1);
                        float4 bs[NC ? NC : 1], go[NC ? NC : 1];
                        if (NC) {
                            ld_px4<NC, NHWC>(base_b, hw, pix, bs);
                            ld_px4<NC, NHWC>(gin_b, hw, pix, go);
                        }
#pragma unroll 3
                        for (int c = 0; c < C; ++c) {
                            float4 bsc, goc;
                            if (NC) { bsc = bs[NC ? c : 0]; goc = go[NC ? c : 0]; }
                            else {
                                bsc = __ldg(reinterpret_cast<const float4*>(base_b + c * hw + pix));
                                goc = __ldg(reinterpret_cast<const float4*>(gin_b + c * hw + pix));
                            }
                            const float bsv[4] = {bsc.x, bsc.y, bsc.z, bsc.w}, gov[4] = {goc.x, goc.y, goc.z, goc.w};
                            float gp[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float pre = bsv[k] + we[k];
                                gp[k] = (pre >= 0.0f && pre <= 1.0f) ? gov[k] : 0.0f;
                                ge[k] = (c == 0) ? gp[k] * wgt : fmaf(gp[k], wgt, ge[k]);
                            }
                            if (NC) go[NC ? c : 0] = make_float4(gp[0], gp[1], gp[2], gp[3]);
                            else if (gbase_b && interior)
                                __stcs(reinterpret_cast<float4*>(gbase_b + c * hw + pix), make_float4(gp[0], gp[1], gp[2], gp[3]));
                        }
                        if (NC && gbase_b && interior) st_px4<NC, NHWC>(gbase_b, hw, pix, go);
                    } else {
                        const float4 t = __ldg(reinterpret_cast<const float4*>(gin_b + pix));
                        ge[0] = t.x; ge[1] = t.y; ge[2] = t.z; ge[3] = t.w;
                    }
                    if (want_gx) {
                        float av[4], bv[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            float gm = g_thin_of_v(variant, a.e.low, a.e.high, mode, ge[k], thin[k], wih[k]);
                            if (meta_removed(meta[k])) gm = 0.0f;                     // core.py:290 / :480
                            // torch.where backward (alpha gate): M == 0 there, and mag_backward returns 0 for mag == 0
                            mag_backward(gm, mag[k], gx1[k], gy1[k], fC, av[k], bv[k]);
                        }
                        st_plane(A + q, av, col == 0, col + 4 == W, 0.0f, 0.0f);
                        st_plane(Bv + q, bv, col == 0, col + 4 == W, 0.0f, 0.0f);
                    }
                }
            }
        }
    }
    if (!want_gx) return;
    __syncthreads();
    float* GB = R2;
    if (active) fast_stage_sobel_adjoint<R, EVEN>(geo, A, Bv, ab_lo, GB, gb_lo, gb_hi, tx, ty);
    __syncthreads();
    if (active) fast_stage_gauss_adjoint_store<NC, R, NHWC, EVEN>(a, geo, GB, gb_lo, a.e.g_x + (size_t)b * C * hw, r0, r1, tx, ty);
}

}  // namespace ee
