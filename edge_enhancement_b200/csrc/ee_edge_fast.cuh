// ee_edge_fast.cuh -- tuned CannyFilter_step125_1 (+ blend) kernels for the 128-bit path.
//
// Same strip decomposition and the same canonical arithmetic as ee_edge_step125.cuh (the generic
// kernels there stay as the any-shape fallback and as a second implementation to test against),
// restructured because the round-1 profile (profiles/r1_*.md) showed the generic kernels are
// ISSUE-bound, not HBM-bound: ~345 (bwd) / 142 (fwd) instructions per pixel, over half of them
// integer / control, and every stencil row recomputed three times.  Changes:
//
//   * shared-memory planes carry one pad column on each side (row stride W+8 floats, data at +4,
//     still 16-byte aligned): replicate / zero extension is materialised by the writer, so a
//     window is always LDS.128 + 2 LDS.32 with no branch or clamp;
//   * register sliding window down the rows: a thread owns one float4 column group and a chunk
//     of R consecutive rows; the horizontal partial sums of each input row (P/Q, D/V, HA/HB) are
//     computed once and reused by the three output rows that need them (3x fewer LDS and a third
//     less FP work); the loop is fully unrolled so the 3-deep ring lives in registers;
//   * x / 3 is evaluated as q = x*r, q' = fma(fma(-q,3,x), r, q) -- the correctly rounded quotient
//     for every finite fp32 x (exhaustively verified), 3 instructions instead of ~10;
//   * the threshold tests never take a square root: sqrt_rn is monotonic, so
//     `sqrt(u) > thr` <=> `u > hi_cut` for a cut-off computed exactly on the host
//     (ee_capi.cu: largest fp32 u with sqrtf(u) <= thr).  sqrt is only evaluated in the backward,
//     on the ~5 % of pixels that carry gradient.
#pragma once
#include <cuda.h>
#include <cuda/barrier>

#include <type_traits>

#include "ee_edge_step125.cuh"

namespace ee {

constexpr int kPadL = 4;      // floats of padding left of column 0 in every plane row (keeps 16B alignment)
constexpr int kPadW = 8;      // total extra floats per plane row

struct FastArgs {
    EdgeArgs e;
    float hi_cut;     // mag >  high   <=> u > hi_cut
    float w_cut;      // mag <= 1.001f <=> u <= w_cut
    float a_cut;      // mag <  alpha  <=> u < a_cut
    float e_cut;      // (mag > high and not mag < alpha) <=> u > e_cut
    float zero_val;   // To_compare's value for "not above": 1 if high < 0 else 0 (core.py:344-345)
    int Wp;           // plane row stride in floats (plane width + 8)
    int TW;           // columns per tile (multiple of 4; == W for a single column tile)
    int tiles_x;      // column tiles per image
    int halo;         // plane halo columns on each side of a column tile (4, or 8 for the Canny backward)
    alignas(64) CUtensorMap x_map;   // TMA-staged kernels only: x as a [B*C, H, W] tensor, box = (plane width + 8) x rows x 1
};

// (sgx, sgy)[4] / C, value-identical to the IEEE division.  DIVM: 0 -> C == 1, 1 -> C == 3, 2 -> any C.
// For C == 3: q = x*r, q' = fma(fma(-q,3,x), r, q) with r = RN(1/3) is the correctly rounded quotient
// for EVERY finite fp32 x including denormals (verified exhaustively over all 2^31 bit patterns
// against x/3.0f; the only difference is the sign of a zero result for x = -0).
template <int DIVM>
__device__ __forceinline__ void div_channels8(const float (&sx)[4], const float (&sy)[4], float fC, float (&gx)[4], float (&gy)[4]) {
    if constexpr (DIVM == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { gx[k] = sx[k]; gy[k] = sy[k]; }
    } else if constexpr (DIVM == 1) {
        const float r = 0.333333343f;                 // RN(1/3) = 0x3eaaaaab
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float qx = sx[k] * r, qy = sy[k] * r;
            gx[k] = fmaf(fmaf(-qx, 3.0f, sx[k]), r, qx);
            gy[k] = fmaf(fmaf(-qy, 3.0f, sy[k]), r, qy);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) { gx[k] = sx[k] / fC; gy[k] = sy[k] / fC; }
    }
}

// edge value from u = gx1^2 + gy1^2 (no sqrt): to_compare(gate(mag), high).  e_cut folds the
// threshold and the alpha gate into one comparison (ee_capi.cu); NaN stays NaN like To_compare.
__device__ __forceinline__ float edge_from_u(const FastArgs& a, float u) {
    return (u > a.e_cut) ? 1.0f : ((u != u) ? u : a.zero_val);
}

// Bulk L2 prefetch (UBLKPF): ask the memory system to start moving a contiguous byte range towards
// L2 now; the LDGs that consume it two stages later then hit L2 instead of waiting on HBM.
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(p), "r"(bytes) : "memory");
}
// rows [lo,hi) of every channel of image b: one contiguous range per channel (or one in total when the
// strip is the whole image).  Issued by the first 2*C threads of the CTA.
__device__ __forceinline__ void prefetch_rows(const float* t, int b, int C, int H, int W, int lo, int hi, int lane) {
    const size_t hw = (size_t)H * W;
    if (lo == 0 && hi == H) {
        if (lane == 0) l2_prefetch_bulk(t + (size_t)b * C * hw, (uint32_t)(C * hw * sizeof(float)));
    } else if (lane < C) {
        l2_prefetch_bulk(t + ((size_t)b * C + lane) * hw + (size_t)lo * W, (uint32_t)((size_t)(hi - lo) * W * sizeof(float)));
    }
}

// torch.clamp(v, 0, 1) with NaN propagation in two instructions (FMNMX.NAN)
__device__ __forceinline__ float clamp01_fast(float v) {
    float r;
    asm("max.NaN.f32 %0, %1, 0f00000000;\n\tmin.NaN.f32 %0, %0, 0f3F800000;" : "=f"(r) : "f"(v));
    return r;
}

// Duplicating every chunk body into a guard-free FULL variant was measured SLOWER on B200 (code
// size / registers: 3.4 vs 5.7 TB/s forward), so it is off by default; kept for experiments.
#ifndef EE_USE_FULL
#define EE_USE_FULL 0
#endif

// L2 bulk prefetch (UBLKPF) of the operands a LATER stage needs (base / g_out).  Placement matters: issued
// before the x loads of the first stage it competes with them (backward -4 %); issued right after them
// (mode 2) the backward gains 10 % (5.0 -> 5.5 TB/s at 4096x3x64x64).  0 = off, 1 = at kernel start, 2 = after
// the first stage's loads.
#ifndef EE_L2_PREFETCH
#define EE_L2_PREFETCH 2
#endif
// (Also tried: prefetching the x rows of the tile that will replace this CTA on its SM, blockIdx + resident
// CTAs.  Measured slower -- forward 6.6 -> 6.3, backward 5.5 -> 5.0 TB/s -- and removed.)
#ifndef EE_L2_PREFETCH_BWD
#define EE_L2_PREFETCH_BWD 2
#endif

// Tile geometry.  When the kernel is specialised on the image width (WT != 0) W, Wp, G and GX are
// compile-time constants, so every row stride becomes an immediate offset and the integer address
// arithmetic (a quarter of the executed instructions in the round-1 profile) disappears.
// A tile is rows [r0,r1) x columns [c0,c1) of one image; its planes cover columns [cs,ce) = the tile plus
// `halo` columns on each side (one float4 group; two for the Canny backward, whose dependency cone is 6
// columns), clipped to the image.  Images up to 128 columns wide use a
// single column tile (cs = 0, ce = W: exactly the full-width strips of round 1); wider images are cut into
// column tiles so that the planes stay small enough for 3 CTAs per SM.
struct Geo { int W, Wp, H, GX, RY, cs, ce, c0, c1, Gt; };
template <int WT, int WG>
__device__ __forceinline__ Geo make_geo(const FastArgs& a, int tile_x) {
    Geo q;
    q.W = WG ? WG : a.e.W;                     // global row stride
    q.Wp = WT ? WT + kPadW : a.Wp;             // plane row stride (WT = plane width incl. halo groups)
    q.H = a.e.H;
    q.GX = WT ? (WT >> 2) : a.e.GX;            // host guarantees GX == plane groups whenever that is <= 256
    q.RY = a.e.RY;
    q.c0 = tile_x * a.TW;
    q.c1 = min(q.c0 + a.TW, q.W);
    q.cs = max(q.c0 - a.halo, 0);
    q.ce = min(q.c1 + a.halo, q.W);
    if (WG && WG == WT) { q.c0 = 0; q.c1 = WG; q.cs = 0; q.ce = WG; }     // single column tile, all constant
    q.Gt = (q.ce - q.cs) >> 2;
    return q;
}

#define EE_FOR_CHUNKS(row_lo, row_hi)                                                         \
    for (int ch = ty, n_ch = ((row_hi) - (row_lo) + R - 1) / R; ch < n_ch; ch += geo.RY)      \
        for (int g = tx; g < geo.Gt; g += geo.GX)
// EVEN (whole-image tile whose height is a multiple of R, one chunk of exactly R rows per active thread):
// the loops run once and every "is this row inside the chunk" guard is compile-time true
#define EE_FOR_CHUNKS_E(row_lo, row_hi)                                                                        \
    for (int ch = ty, n_ch = ((row_hi) - (row_lo) + R - 1) / R, o1_ = 1; EVEN ? o1_ : ch < n_ch;               \
         ch += geo.RY, o1_ = 0)                                                                                \
        for (int g = tx, o2_ = 1; EVEN ? o2_ : g < geo.Gt; g += geo.GX, o2_ = 0)

// Every chunk body exists twice: FULL (all R rows present, no image border inside the chunk: no
// guards, no clamps) and the guarded general version.  `full_t` / `part_t` select them.
using full_t = std::integral_constant<bool, true>;
using part_t = std::integral_constant<bool, false>;

__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// write 4 values + the pad column a border group owns
__device__ __forceinline__ void st_plane(float* q, const float (&o)[4], bool left, bool right, float lpad, float rpad) {
    *reinterpret_cast<float4*>(q) = make_float4(o[0], o[1], o[2], o[3]);
    if (left) q[-1] = lpad;
    if (right) q[4] = rpad;
}

// ---- channel I/O of 4 consecutive pixels -----------------------------------------------------------------
// v[c] holds channel c of pixels pix..pix+3 of ONE image (img = that image's first float; an image is C*hw
// contiguous floats in both layouts).  NCHW: one 128-bit access per channel plane.  NHWC (channels_last,
// C = 3): the 12 floats of the 4 pixels are 3 consecutive 128-bit words that are (de)interleaved in registers.
template <int NC, bool NHWC>
__device__ __forceinline__ void ld_px4(const float* __restrict__ img, size_t hw, int pix, float4 (&v)[NC ? NC : 1]) {
    if constexpr (NHWC) {
        static_assert(NC == 3, "channels_last path is specialised for C = 3");
        const float4* p = reinterpret_cast<const float4*>(img + (size_t)pix * 3);
        const float4 t0 = __ldg(p), t1 = __ldg(p + 1), t2 = __ldg(p + 2);
        v[0] = make_float4(t0.x, t0.w, t1.z, t2.y);
        v[1] = make_float4(t0.y, t1.x, t1.w, t2.z);
        v[2] = make_float4(t0.z, t1.y, t2.x, t2.w);
    } else {
#pragma unroll
        for (int c = 0; c < (NC ? NC : 1); ++c) v[c] = __ldg(reinterpret_cast<const float4*>(img + c * hw + pix));
    }
}
template <int NC, bool NHWC>
__device__ __forceinline__ void st_px4(float* img, size_t hw, int pix, const float4 (&v)[NC ? NC : 1]) {
    if constexpr (NHWC) {
        static_assert(NC == 3, "channels_last path is specialised for C = 3");
        float4* p = reinterpret_cast<float4*>(img + (size_t)pix * 3);
        __stcs(p, make_float4(v[0].x, v[1].x, v[2].x, v[0].y));
        __stcs(p + 1, make_float4(v[1].y, v[2].y, v[0].z, v[1].z));
        __stcs(p + 2, make_float4(v[2].z, v[0].w, v[1].w, v[2].w));
    } else {
#pragma unroll
        for (int c = 0; c < (NC ? NC : 1); ++c) __stcs(reinterpret_cast<float4*>(img + c * hw + pix), v[c]);
    }
}

// ---- stage S: channel sum of x rows [lo,hi) into a replicate-padded plane -------------------
template <int NC, int R, bool NHWC = false, bool EVEN = false>
__device__ __forceinline__ void fast_stage_sum(const FastArgs& a, const Geo geo, const float* __restrict__ xb, float* S,
                                               int lo, int hi, int tx, int ty) {
    const int W = geo.W, Wp = geo.Wp;
    const int C = NC ? NC : a.e.C;
    const size_t hw = (size_t)geo.H * W;
    EE_FOR_CHUNKS_E(lo, hi) {
        const int lc = g * 4, col = geo.cs + lc, ra = lo + ch * R;
        const float* px = xb + (size_t)ra * W + col;
        float* ps = S + (size_t)(ra - lo) * Wp + kPadL + lc;
        auto body = [&](auto tag) {
            constexpr bool FULL = decltype(tag)::value || EVEN;
            float4 acc[R];
            if (NHWC) {
                float4 v1[R], v2[R];
                const float4* p3 = reinterpret_cast<const float4*>(xb + ((size_t)ra * W + col) * 3);
#pragma unroll
                for (int i = 0; i < R; ++i)
                    if (FULL || ra + i < hi) {
                        acc[i] = __ldg(p3 + (size_t)i * (3 * W / 4));
                        v1[i] = __ldg(p3 + (size_t)i * (3 * W / 4) + 1);
                        v2[i] = __ldg(p3 + (size_t)i * (3 * W / 4) + 2);
                    }
#pragma unroll
                for (int i = 0; i < R; ++i)
                    if (FULL || ra + i < hi)      // same ((c0 + c1) + c2) order as the planar path
                        acc[i] = make_float4((acc[i].x + acc[i].y) + acc[i].z, (acc[i].w + v1[i].x) + v1[i].y,
                                             (v1[i].z + v1[i].w) + v2[i].x, (v2[i].y + v2[i].z) + v2[i].w);
            } else if (NC == 3) {
                float4 v1[R], v2[R];
#pragma unroll
                for (int i = 0; i < R; ++i)
                    if (FULL || ra + i < hi) {
                        acc[i] = __ldg(reinterpret_cast<const float4*>(px + i * W));
                        v1[i] = __ldg(reinterpret_cast<const float4*>(px + i * W + hw));
                        v2[i] = __ldg(reinterpret_cast<const float4*>(px + i * W + 2 * hw));
                    }
#pragma unroll
                for (int i = 0; i < R; ++i)
                    if (FULL || ra + i < hi) acc[i] = f4add(f4add(acc[i], v1[i]), v2[i]);
            } else {
#pragma unroll
                for (int i = 0; i < R; ++i)
                    if (FULL || ra + i < hi) acc[i] = __ldg(reinterpret_cast<const float4*>(px + i * W));
                for (int c = 1; c < C; ++c) {
#pragma unroll
                    for (int i = 0; i < R; ++i)
                        if (FULL || ra + i < hi) acc[i] = f4add(acc[i], __ldg(reinterpret_cast<const float4*>(px + i * W + (size_t)c * hw)));
                }
            }
#pragma unroll
            for (int i = 0; i < R; ++i)
                if (FULL || ra + i < hi) {
                    const float o[4] = {acc[i].x, acc[i].y, acc[i].z, acc[i].w};
                    st_plane(ps + i * Wp, o, col == 0, col + 4 == W, o[0], o[3]);
                }
        };
        if (EE_USE_FULL && ra + R <= hi) body(full_t{}); else body(part_t{});
    }
}

// window of one plane row: l | m[4] | r
struct Win { float l, m0, m1, m2, m3, r; };
__device__ __forceinline__ Win ld_win(const float* p) {
    const float4 m = *reinterpret_cast<const float4*>(p);
    Win w; w.l = p[-1]; w.m0 = m.x; w.m1 = m.y; w.m2 = m.z; w.m3 = m.w; w.r = p[4];
    return w;
}

// Gaussian horizontal partials of one row: P (outer rows) and Q (centre row)
__device__ __forceinline__ void gauss_partials(const Win& w, float c0, float c1, float c2, float (&P)[4], float (&Q)[4]) {
    const float e0 = w.l + w.m1, e1 = w.m0 + w.m2, e2 = w.m1 + w.m3, e3 = w.m2 + w.r;
    P[0] = fmaf(c1, w.m0, c0 * e0); P[1] = fmaf(c1, w.m1, c0 * e1); P[2] = fmaf(c1, w.m2, c0 * e2); P[3] = fmaf(c1, w.m3, c0 * e3);
    Q[0] = fmaf(c2, w.m0, c1 * e0); Q[1] = fmaf(c2, w.m1, c1 * e1); Q[2] = fmaf(c2, w.m2, c1 * e2); Q[3] = fmaf(c2, w.m3, c1 * e3);
}

// Forward sliding chunks: output rows [ra, rb), inputs ra-1 .. rb with the row index clamped to the
// image (replicate padding).  A chunk is FULL when it has R rows and touches neither image border.

#define EE_FWD_CHUNK_IS_FULL(ra, rb, H) (EE_USE_FULL && (rb) - (ra) == R && (ra) > 0 && (rb) < (H))

// ---- stage blur: Bl rows [lo,hi) from S ------------------------------------------------------
template <int R, bool EVEN = false>
__device__ __forceinline__ void fast_stage_blur(const FastArgs& a, const Geo geo, const float* S, int s_lo, float* Bl,
                                                int lo, int hi, int tx, int ty) {
    const int W = geo.W, H = geo.H, Wp = geo.Wp;
    const float c0 = a.e.c0, c1 = a.e.c1, c2 = a.e.c2;
    EE_FOR_CHUNKS_E(lo, hi) {
        const int lc = g * 4, col = geo.cs + lc, ra = lo + ch * R, rb = EVEN ? ra + R : min(ra + R, hi);
        const float* ps = S + kPadL + lc;                 // row r at ps + (r - s_lo) * Wp
        float* pb = Bl + (size_t)(ra - lo) * Wp + kPadL + lc;
        auto body = [&](auto tag) {
            constexpr bool FULL = decltype(tag)::value;
            float P[3][4], Q[3][4];
#pragma unroll
            for (int i = 0; i < R + 2; ++i) {
                const int rin = ra - 1 + i;
                if (FULL || EVEN || rin <= rb) {
                    const int rc = FULL ? rin : min(max(rin, 0), H - 1);
                    gauss_partials(ld_win(ps + (rc - s_lo) * Wp), c0, c1, c2, P[i % 3], Q[i % 3]);
                }
                if (i >= 2 && (FULL || EVEN || ra + i - 2 < rb)) {
                    float o[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) o[k] = (P[(i - 2) % 3][k] + Q[(i - 1) % 3][k]) + P[i % 3][k];
                    st_plane(pb + (i - 2) * Wp, o, col == 0, col + 4 == W, o[0], o[3]);
                }
            }
        };
        if (EE_FWD_CHUNK_IS_FULL(ra, rb, H)) body(full_t{}); else body(part_t{});
    }
}

// Sobel horizontal partials of one blurred row: D = right - left, V = fma(.5, left + right, mid)
__device__ __forceinline__ void sobel_partials(const Win& w, float (&D)[4], float (&V)[4]) {
    D[0] = w.m1 - w.l; D[1] = w.m2 - w.m0; D[2] = w.m3 - w.m1; D[3] = w.r - w.m2;
    V[0] = fmaf(0.5f, w.l + w.m1, w.m0); V[1] = fmaf(0.5f, w.m0 + w.m2, w.m1);
    V[2] = fmaf(0.5f, w.m1 + w.m3, w.m2); V[3] = fmaf(0.5f, w.m2 + w.r, w.m3);
}

// -------------------------------------------------------------------------------------------
// forward:  planes S (TH+4 rows) and Bl (TH+2 rows), both with stride Wp
// -------------------------------------------------------------------------------------------
// HT != 0: the tile is the whole image, H == HT is a compile-time constant and a multiple of R, so every active
// thread owns exactly one chunk of R rows in every stage ("EVEN"): chunk loops and row guards disappear.
template <int NC, bool BLEND, int R, int WT, int WG, bool NHWC = false, int HT = 0>
#ifndef EE_MINB_FWD
#define EE_MINB_FWD 3
#endif
#ifndef EE_MINB_FWD_EVEN
#define EE_MINB_FWD_EVEN 4
#endif
#ifndef EE_MINB_BWD
#define EE_MINB_BWD 3
#endif
// whole-image tiles: 64 registers / 4 CTAs per SM measured +3 % (6.47 -> 6.69 TB/s at 64 px); strips keep 80 registers / 3
__global__ void __launch_bounds__(256, HT ? EE_MINB_FWD_EVEN : EE_MINB_FWD) edge_fwd_step125_fast(const FastArgs a) {
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

    const int s_lo = max(r0 - 2, 0), s_hi = min(r1 + 2, H);
    const int b_lo = max(r0 - 1, 0), b_hi = min(r1 + 1, H);
    float* S = smem;
    float* Bl = smem + (size_t)(EVEN ? HT : min(a.e.TH + 4, H)) * Wp;
    const bool one_tile = EVEN || a.tiles_x == 1;

#if EE_L2_PREFETCH == 1
    if (BLEND && (!NHWC || (r0 == 0 && r1 == H)) && one_tile && threadIdx.x < 32 && C <= 32) prefetch_rows(a.e.base, b, C, H, W, r0, r1, threadIdx.x);
#endif
    if (ty < geo.RY) fast_stage_sum<NC, R, NHWC, EVEN>(a, geo, a.e.x + (size_t)b * C * hw, S, s_lo, s_hi, tx, ty);
#if EE_L2_PREFETCH == 2
    if (BLEND && (!NHWC || (r0 == 0 && r1 == H)) && one_tile && threadIdx.x < 32 && C <= 32) prefetch_rows(a.e.base, b, C, H, W, r0, r1, threadIdx.x);
#endif
    __syncthreads();
    if (ty < geo.RY) fast_stage_blur<R, EVEN>(a, geo, S, s_lo, Bl, b_lo, b_hi, tx, ty);
    __syncthreads();
    if (ty >= geo.RY) return;

    const float fC = a.e.fC, wgt = a.e.w;
    const float* base_b = a.e.base + (size_t)b * C * hw;
    float* out_b = a.e.out + (size_t)b * C * hw;
    EE_FOR_CHUNKS_E(r0, r1) {
        const int lc = g * 4, col = geo.cs + lc, ra = r0 + ch * R, rb = EVEN ? ra + R : min(ra + R, r1);
        if (!EVEN && (col < geo.c0 || col >= geo.c1)) continue;  // halo groups produce no output
        const float* pbl = Bl + kPadL + lc;
        auto body = [&](auto tag) {
            constexpr bool FULL = decltype(tag)::value;
            float D[3][4], V[3][4];
#pragma unroll
            for (int i = 0; i < R + 2; ++i) {
                const int rin = ra - 1 + i;
                if (FULL || EVEN || rin <= rb) {
                    const int rc = FULL ? rin : min(max(rin, 0), H - 1);
                    sobel_partials(ld_win(pbl + (rc - b_lo) * Wp), D[i % 3], V[i % 3]);
                }
                if (i >= 2 && (FULL || EVEN || ra + i - 2 < rb)) {
                    const int pix = (ra + i - 2) * W + col;
                    float4 bs[NC ? NC : 1];
                    if (BLEND && NC) ld_px4<NC, NHWC>(base_b, hw, pix, bs);
                    float e[4], sgx[4], sgy[4], gx1[4], gy1[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        sgx[k] = fmaf(0.5f, D[(i - 2) % 3][k] + D[i % 3][k], D[(i - 1) % 3][k]);
                        sgy[k] = V[i % 3][k] - V[(i - 2) % 3][k];
                    }
                    div_channels8<DIVM>(sgx, sgy, fC, gx1, gy1);
#pragma unroll
                    for (int k = 0; k < 4; ++k) e[k] = edge_from_u(a, gx1[k] * gx1[k] + gy1[k] * gy1[k]);
                    if (a.e.edge) __stcs(reinterpret_cast<float4*>(a.e.edge + (size_t)b * hw + pix), make_float4(e[0], e[1], e[2], e[3]));
                    if (BLEND) {
                        const float w0 = wgt * e[0], w1 = wgt * e[1], w2 = wgt * e[2], w3 = wgt * e[3];
                        if (NC) {
#pragma unroll
                            for (int c = 0; c < (NC ? NC : 1); ++c)
                                bs[c] = make_float4(clamp01_fast(bs[c].x + w0), clamp01_fast(bs[c].y + w1),
                                                    clamp01_fast(bs[c].z + w2), clamp01_fast(bs[c].w + w3));
                            st_px4<NC, NHWC>(out_b, hw, pix, bs);
                        } else {
                            for (int c = 0; c < C; ++c) {
                                const float4 t = __ldg(reinterpret_cast<const float4*>(base_b + c * hw + pix));
                                const float4 o = make_float4(clamp01_fast(t.x + w0), clamp01_fast(t.y + w1),
                                                             clamp01_fast(t.z + w2), clamp01_fast(t.w + w3));
                                __stcs(reinterpret_cast<float4*>(out_b + c * hw + pix), o);
                            }
                        }
                    }
                }
            }
        };
        if (EE_FWD_CHUNK_IS_FULL(ra, rb, H)) body(full_t{}); else body(part_t{});
    }
}

// -------------------------------------------------------------------------------------------
// adjoint sliding stages.  Output rows p run over [pa, pb) in padded-frame coordinates: the chunk
// that owns image row 0 also produces ring row -1 and folds it into row 0; the chunk that owns row
// H-1 also produces ring row H and folds it into row H-1.  Ring columns -1 / W are produced by the
// border groups (one scalar per row) and folded into columns 0 / W-1.  Zero extension: input rows
// outside the image contribute zero partials.  A chunk is FULL when it has R rows and touches
// neither image border: then there are no ring rows and no bounds checks.
// -------------------------------------------------------------------------------------------
struct AdjBorder { bool left, right; };

// Sobel-adjoint partials of input row (a-row wa, b-row wb): HA = a(q-1) - a(q+1), HB = fma(.5, b(q-1)+b(q+1), b(q)),
// plus the ring column partials for a border group.
__device__ __forceinline__ void sobel_adj_partials(const Win& wa, const Win& wb, AdjBorder bd, float (&HA)[4], float (&HB)[4],
                                                   float& HAr, float& HBr) {
    HA[0] = wa.l - wa.m1; HA[1] = wa.m0 - wa.m2; HA[2] = wa.m1 - wa.m3; HA[3] = wa.m2 - wa.r;
    HB[0] = fmaf(0.5f, wb.l + wb.m1, wb.m0); HB[1] = fmaf(0.5f, wb.m0 + wb.m2, wb.m1);
    HB[2] = fmaf(0.5f, wb.m1 + wb.m3, wb.m2); HB[3] = fmaf(0.5f, wb.m2 + wb.r, wb.m3);
    // ring column: left  q=-1: HA = 0 - a(0),   HB = fma(.5, 0 + b(0), 0)
    //              right q=W : HA = a(W-1) - 0, HB = fma(.5, b(W-1) + 0, 0)
    const float alo = bd.left ? 0.0f : wa.m3, ahi = bd.left ? wa.m0 : 0.0f;
    HAr = alo - ahi;
    HBr = fmaf(0.5f, 0.0f + (bd.left ? wb.m0 : wb.m3), 0.0f);
}

__device__ __forceinline__ void gauss_adj_partials(const Win& w, AdjBorder bd, float c0, float c1, float c2, float (&P)[4],
                                                   float (&Q)[4], float& Pr, float& Qr) {
    gauss_partials(w, c0, c1, c2, P, Q);
    // ring column: e = g(q-1) + g(q+1) with one of them 0 and m = g(q) = 0
    const float er = 0.0f + (bd.left ? w.m0 : w.m3);
    Pr = fmaf(c1, 0.0f, c0 * er);
    Qr = fmaf(c2, 0.0f, c1 * er);
}

// Generic adjoint sliding chunk.  LOADP(i, rin, valid) fills partial slot i%3 (zeros when !valid),
// COMBINE(i, o) combines slots (i-2, i-1, i) into o[4] incl. the ring column, STORE(row, o) emits.
template <int R, bool FULL, typename LoadP, typename Combine, typename Store>
__device__ __forceinline__ void adj_chunk(int ra, int rb, int H, LoadP loadp, Combine combine, Store store) {
    if constexpr (FULL) {
#pragma unroll
        for (int i = 0; i < R + 2; ++i) {
            loadp(i, ra - 1 + i, true);
            if (i >= 2) {
                float o[4];
                combine(i, o);
                store(ra + i - 2, o);
            }
        }
    } else {
        const int pa = ra - (ra == 0 ? 1 : 0), pb = rb + (rb == H ? 1 : 0);      // output rows incl. ring rows
        float hold[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int i = 0; i < R + 4; ++i) {
            const int rin = ra - 2 + i;
            if (rin >= pa - 1 && rin <= pb) loadp(i, rin, rin >= 0 && rin < H);
            if (i >= 2) {
                const int p = ra - 3 + i;                     // centre of the last three inputs
                if (p >= pa && p < pb) {
                    float o[4];
                    combine(i, o);
                    if (p == -1) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) hold[k] = o[k];                       // ring row above, folded into row 0
                    } else if (p == H) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) o[k] = hold[k] + o[k];                // row H-1 + ring row below
                        store(H - 1, o);
                    } else {
                        if (p == 0) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) o[k] = o[k] + hold[k];
                        }
                        if (p == H - 1) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) hold[k] = o[k];                   // wait for ring row H
                        } else {
                            store(p, o);
                        }
                    }
                }
            }
        }
    }
}

// EVEN chunk of a whole-image tile: exactly R output rows [ra, ra+R), ra % R == 0, H % R == 0.  Inputs ra-1 .. ra+R
// (zero partials outside the image); only the chunks with ra == 0 / ra + R == H produce and fold a ring row.  Ring row
// -1 combines (zero, zero, row 0), ring row H combines (row H-1, zero, zero): the ring-buffer slot that is not loaded
// yet (or no longer needed) is zeroed and `combine` is called with the index whose slots line up.
template <int R, typename LoadP, typename Combine, typename Store>
__device__ __forceinline__ void adj_chunk_even(int ra, int H, LoadP loadp, Combine combine, Store store) {
    const bool top = (ra == 0), bot = (ra + R == H);
    float hold[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    loadp(0, ra - 1, !top);
    loadp(1, ra, true);
    if (top) {
        loadp(2, -2, false);
        combine(4, hold);                 // slots (2, 0, 1) = (zero, zero, row 0): ring row -1
    }
#pragma unroll
    for (int i = 2; i < R + 2; ++i) {
        loadp(i, ra - 1 + i, (i < R + 1) || !bot);
        float o[4];
        combine(i, o);
        if (i == 2 && top) {
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = o[k] + hold[k];
        }
        if (i == R + 1 && bot) {
            float ring[4];
            loadp(R + 2, H + 1, false);
            combine(R + 2, ring);         // slots (R, R+1, R+2) = (row H-1, zero, zero): ring row H
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = o[k] + ring[k];
        }
        store(ra + i - 2, o);
    }
}

// GB rows [gb_lo, gb_hi) = fold(Sobel^T(A, Bv)), zero pad columns.  A / Bv planes start at row ab_lo.
template <int R, bool EVEN = false>
__device__ __forceinline__ void fast_stage_sobel_adjoint(const Geo geo, const float* A, const float* Bv, int ab_lo,
                                                         float* GB, int gb_lo, int gb_hi, int tx, int ty) {
    const int W = geo.W, H = geo.H, Wp = geo.Wp;
    EE_FOR_CHUNKS_E(gb_lo, gb_hi) {
        const int lc = g * 4, col = geo.cs + lc, ra = gb_lo + ch * R, rb = EVEN ? ra + R : min(ra + R, gb_hi);
        const AdjBorder bd = {col == 0, col + 4 == W};
        const bool ring = bd.left || bd.right;
        float HA[3][4], HB[3][4], HAr[3], HBr[3];
        const float* pA = A + kPadL + lc;
        const float* pB = Bv + kPadL + lc;
        auto loadp = [&](int i, int rin, bool valid) {
            if (valid) {
                const int q = (rin - ab_lo) * Wp;
                sobel_adj_partials(ld_win(pA + q), ld_win(pB + q), bd, HA[i % 3], HB[i % 3], HAr[i % 3], HBr[i % 3]);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) { HA[i % 3][k] = 0.0f; HB[i % 3][k] = 0.0f; }
                HAr[i % 3] = 0.0f; HBr[i % 3] = 0.0f;
            }
        };
        auto combine = [&](int i, float (&o)[4]) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float xa = fmaf(0.5f, HA[(i - 2) % 3][k] + HA[i % 3][k], HA[(i - 1) % 3][k]);
                const float yb = HB[(i - 2) % 3][k] - HB[i % 3][k];
                o[k] = xa + yb;
            }
            if (ring) {
                const float xa = fmaf(0.5f, HAr[(i - 2) % 3] + HAr[i % 3], HAr[(i - 1) % 3]);
                const float yb = HBr[(i - 2) % 3] - HBr[i % 3];
                const float t = xa + yb;
                if (bd.left) o[0] = o[0] + t; else o[3] = o[3] + t;
            }
        };
        auto store = [&](int row, const float (&o)[4]) {
            st_plane(GB + (row - gb_lo) * Wp + kPadL + lc, o, bd.left, bd.right, 0.0f, 0.0f);
        };
        if constexpr (EVEN) adj_chunk_even<R>(ra, H, loadp, combine, store);
        else if (EE_FWD_CHUNK_IS_FULL(ra, rb, H)) adj_chunk<R, true>(ra, rb, H, loadp, combine, store);
        else adj_chunk<R, false>(ra, rb, H, loadp, combine, store);
    }
}

// g_s rows [r0, r1) = fold(Gauss^T(GB)), written to every channel of g_x (gx_b = image base pointer)
template <int NC, int R, bool NHWC = false, bool EVEN = false>
__device__ __forceinline__ void fast_stage_gauss_adjoint_store(const FastArgs& a, const Geo geo, const float* GB, int gb_lo,
                                                               float* gx_b, int r0, int r1, int tx, int ty) {
    const int W = geo.W, H = geo.H, Wp = geo.Wp;
    const int C = NC ? NC : a.e.C;
    const size_t hw = (size_t)H * W;
    const float c0 = a.e.c0, c1 = a.e.c1, c2 = a.e.c2;
    EE_FOR_CHUNKS_E(r0, r1) {
        const int lc = g * 4, col = geo.cs + lc, ra = r0 + ch * R, rb = EVEN ? ra + R : min(ra + R, r1);
        if (!EVEN && (col < geo.c0 || col >= geo.c1)) continue;  // halo groups produce no output
        const AdjBorder bd = {col == 0, col + 4 == W};
        const bool ring = bd.left || bd.right;
        float P[3][4], Q[3][4], Pr[3], Qr[3];
        const float* pG = GB + kPadL + lc;
        auto loadp = [&](int i, int rin, bool valid) {
            if (valid) {
                gauss_adj_partials(ld_win(pG + (rin - gb_lo) * Wp), bd, c0, c1, c2, P[i % 3], Q[i % 3], Pr[i % 3], Qr[i % 3]);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) { P[i % 3][k] = 0.0f; Q[i % 3][k] = 0.0f; }
                Pr[i % 3] = 0.0f; Qr[i % 3] = 0.0f;
            }
        };
        auto combine = [&](int i, float (&o)[4]) {
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = (P[(i - 2) % 3][k] + Q[(i - 1) % 3][k]) + P[i % 3][k];
            if (ring) {
                const float t = (Pr[(i - 2) % 3] + Qr[(i - 1) % 3]) + Pr[i % 3];
                if (bd.left) o[0] = o[0] + t; else o[3] = o[3] + t;
            }
        };
        auto store = [&](int row, const float (&o)[4]) {
            const float4 v = make_float4(o[0], o[1], o[2], o[3]);
            if (NC) {
                float4 vv[NC ? NC : 1];
#pragma unroll
                for (int c = 0; c < (NC ? NC : 1); ++c) vv[c] = v;
                st_px4<NC, NHWC>(gx_b, hw, row * W + col, vv);
            } else {
                float* pg = gx_b + row * W + col;
                for (int c = 0; c < C; ++c) __stcs(reinterpret_cast<float4*>(pg + c * hw), v);
            }
        };
        if constexpr (EVEN) adj_chunk_even<R>(ra, H, loadp, combine, store);
        else if (EE_FWD_CHUNK_IS_FULL(ra, rb, H)) adj_chunk<R, true>(ra, rb, H, loadp, combine, store);
        else adj_chunk<R, false>(ra, rb, H, loadp, combine, store);
    }
}

// One output row (4 pixels) of the backward's A/Bv stage: from the Sobel partials of the three blurred rows around it
// recompute gx1, gy1, u = mag^2 and the edge value, apply the blend's clamp mask to g_out (-> g_base, written when
// `interior`), reduce over channels, apply the STE window and the magnitude adjoint, and store A = dL/dSgx,
// Bv = dL/dSgy into the planes (pad columns zeroed by the border groups).
template <int NC, bool BLEND, bool NHWC, int DIVM>
__device__ __forceinline__ void bwd_abv_row(const FastArgs& a, const float (&Du)[4], const float (&Dm)[4], const float (&Dd)[4],
                                            const float (&Vu)[4], const float (&Vd)[4], int pix, bool interior, bool left,
                                            bool right, const float* __restrict__ base_b, const float* __restrict__ gin_b,
                                            float* gbase_b, size_t hw, bool want_gx, float* Adst, float* Bvdst) {
    const int C = NC ? NC : a.e.C;
    const float fC = a.e.fC, wgt = a.e.w;
    float gx1[4], gy1[4], u[4], ge[4], sgx[4], sgy[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        sgx[k] = fmaf(0.5f, Du[k] + Dd[k], Dm[k]);
        sgy[k] = Vd[k] - Vu[k];
    }
    div_channels8<DIVM>(sgx, sgy, fC, gx1, gy1);
#pragma unroll
    for (int k = 0; k < 4; ++k) u[k] = gx1[k] * gx1[k] + gy1[k] * gy1[k];
    if (BLEND) {
        float we[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) we[k] = wgt * edge_from_u(a, u[k]);
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
            if (NC) go[NC ? c : 0] = make_float4(gp[0], gp[1], gp[2], gp[3]);     // g_base of this channel
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
        bool any = false;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            // To_compare.backward window and the torch.where gate in u-space:
            // (mag > high, mag <= 1.001, not mag < alpha) <=> e_cut < u <= w_cut
            const bool in_win = (u[k] > a.e_cut) && (u[k] <= a.w_cut);
            ge[k] = in_win ? ge[k] : 0.0f;
            av[k] = 0.0f; bv[k] = 0.0f;
            any = any || (ge[k] != 0.0f);
        }
        if (any) {          // ~5 % of pixels carry gradient: one branch per 4 pixels
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (ge[k] != 0.0f && u[k] != 0.0f) {
                    const float t = ge[k] / (sqrtf(u[k]) * fC);
                    av[k] = t * gx1[k];
                    bv[k] = t * gy1[k];
                }
            }
        }
        st_plane(Adst, av, left, right, 0.0f, 0.0f);
        st_plane(Bvdst, bv, left, right, 0.0f, 0.0f);
    }
}

// -------------------------------------------------------------------------------------------
// backward.  smem regions (stride Wp): R1 = S then A (TH+8 rows), R2 = Bl then GB (TH+6), R3 = Bv (TH+4);
// a region never needs more rows than the image has (halo rows are clipped), so each is min(TH+k, H) rows
// -------------------------------------------------------------------------------------------
// TMA (whole-image tiles, C = 3, NCHW): the three channel planes of x -- INCLUDING the pad columns, i.e. a box of W + 8
// columns x H rows starting at column -4 (out-of-bounds columns zero-filled) -- are staged by the TMA engine
// (cp.async.bulk.tensor.3d, one elected thread, mbarrier completion) straight into the three plane regions, which are
// all dead at kernel start; the channel sum then runs shared -> shared in place.  Otherwise 128-bit LDGs are summed in
// registers.
template <int NC, bool BLEND, int R, int WT, int WG, bool NHWC = false, int HT = 0, bool TMA = false>
__global__ void __launch_bounds__(256, EE_MINB_BWD) edge_bwd_step125_fast(const __grid_constant__ FastArgs a) {
    extern __shared__ __align__(128) float smem_bwd125[];        // TMA destinations must be 128-byte aligned
    float* smem = smem_bwd125;
    static_assert(!TMA || (HT != 0 && NC == 3 && !NHWC), "TMA staging: whole-image tiles, C = 3, NCHW");
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
    const bool one_tile = EVEN || a.tiles_x == 1;

    float* R1 = smem;
    float* R2 = R1 + (size_t)(EVEN ? HT : min(a.e.TH + 8, H)) * Wp;
    float* R3 = R2 + (size_t)(EVEN ? HT : min(a.e.TH + 6, H)) * Wp;

    const bool want_gx = (a.e.g_x != nullptr);
    const int s_lo = max(r0 - 4, 0), s_hi = min(r1 + 4, H);
    const int b_lo = max(r0 - 3, 0), b_hi = min(r1 + 3, H);
    const int ab_lo = want_gx ? max(r0 - 2, 0) : r0, ab_hi = want_gx ? min(r1 + 2, H) : r1;
    const int gb_lo = max(r0 - 1, 0), gb_hi = min(r1 + 1, H);

    float* S = R1; float* Bl = R2;
    auto prefetch_bwd_operands = [&]() {
        if (NHWC && !(ab_lo == 0 && ab_hi == H)) return;     // channels_last: only the whole-image range is contiguous the same way
        if (C <= 32 && one_tile) {
            if (BLEND) {
                if (threadIdx.x < 32) prefetch_rows(a.e.base, b, C, H, W, ab_lo, ab_hi, threadIdx.x);
                else if (threadIdx.x < 64) prefetch_rows(a.e.g_in, b, C, H, W, ab_lo, ab_hi, threadIdx.x - 32);
            } else if (threadIdx.x == 0) {
                prefetch_rows(a.e.g_in, b, 1, H, W, ab_lo, ab_hi, 0);
            }
        } else if (BLEND && !one_tile) {
            // column tile: one bulk prefetch per (tensor, channel, row) segment [cs, ce) of the A/Bv rows
            const int nrows = ab_hi - ab_lo, seg = (geo.ce - geo.cs) * (int)sizeof(float);
            for (int i = threadIdx.x; i < 2 * C * nrows; i += blockDim.x) {
                const int t = i / (C * nrows), rem = i - t * (C * nrows), c = rem / nrows, r = ab_lo + rem - c * nrows;
                const float* src = (t ? a.e.g_in : a.e.base) + ((size_t)b * C + c) * hw + (size_t)r * W + geo.cs;
                l2_prefetch_bulk(src, (uint32_t)seg);
            }
        }
    };
#if EE_L2_PREFETCH_BWD == 1
    prefetch_bwd_operands();
#endif
    if constexpr (TMA) {
        namespace cde = cuda::device::experimental;
#pragma nv_diag_suppress static_var_with_dynamic_init
        __shared__ cuda::barrier<cuda::thread_scope_block> bar;
        if (threadIdx.x == 0) {
            init(&bar, blockDim.x);
            cde::fence_proxy_async_shared_cta();
        }
        __syncthreads();
        cuda::barrier<cuda::thread_scope_block>::arrival_token tok;
        if (threadIdx.x == 0) {
            cde::cp_async_bulk_tensor_3d_global_to_shared(R1, &a.x_map, -kPadL, 0, b * 3 + 0, bar);
            cde::cp_async_bulk_tensor_3d_global_to_shared(R2, &a.x_map, -kPadL, 0, b * 3 + 1, bar);
            cde::cp_async_bulk_tensor_3d_global_to_shared(R3, &a.x_map, -kPadL, 0, b * 3 + 2, bar);
            tok = cuda::device::barrier_arrive_tx(bar, 1, 3u * (uint32_t)(HT * (WT + kPadW) * sizeof(float)));
        } else {
            tok = bar.arrive();
        }
        prefetch_bwd_operands();    // the x tiles are in flight: ask for base / g_out right behind them
        bar.wait(std::move(tok));
        if (active) {
            const int q0 = ty * R * Wp + kPadL + tx * 4;
#pragma unroll
            for (int i = 0; i < R; ++i) {
                const int q = q0 + i * Wp;
                const float4 v0 = *reinterpret_cast<const float4*>(R1 + q), v1 = *reinterpret_cast<const float4*>(R2 + q),
                             v2 = *reinterpret_cast<const float4*>(R3 + q);
                const float4 sum = f4add(f4add(v0, v1), v2);
                const float o[4] = {sum.x, sum.y, sum.z, sum.w};
                st_plane(S + q, o, tx == 0, tx == geo.GX - 1, o[0], o[3]);
            }
        }
    } else {
        if (active) fast_stage_sum<NC, R, NHWC, EVEN>(a, geo, a.e.x + (size_t)b * C * hw, S, s_lo, s_hi, tx, ty);
#if EE_L2_PREFETCH_BWD == 2
        prefetch_bwd_operands();        // after the x loads are issued, so they do not compete with them
#endif
    }
    __syncthreads();
    if (active) fast_stage_blur<R, EVEN>(a, geo, S, s_lo, Bl, b_lo, b_hi, tx, ty);
    __syncthreads();

    // ---- A / Bv = dL/dSgx, dL/dSgy on rows [ab_lo, ab_hi), zero pad columns ---------------------
    float* A = R1; float* Bv = R3;
    if (active) {
        const float* base_b = a.e.base + (size_t)b * C * hw;
        const float* gin_b = a.e.g_in + (size_t)b * (BLEND ? C : 1) * hw;
        float* gbase_b = a.e.g_base ? a.e.g_base + (size_t)b * C * hw : nullptr;
        EE_FOR_CHUNKS_E(ab_lo, ab_hi) {
            const int lc = g * 4, col = geo.cs + lc, ra = ab_lo + ch * R, rb = EVEN ? ra + R : min(ra + R, ab_hi);
            const float* pbl = Bl + kPadL + lc;
            auto body = [&](auto tag) {
                constexpr bool FULL = decltype(tag)::value;
                float D[3][4], V[3][4];
#pragma unroll
                for (int i = 0; i < R + 2; ++i) {
                    const int rin = ra - 1 + i;
                    if (FULL || EVEN || rin <= rb) {
                        const int rc = FULL ? rin : min(max(rin, 0), H - 1);
                        sobel_partials(ld_win(pbl + (rc - b_lo) * Wp), D[i % 3], V[i % 3]);
                    }
                    if (i >= 2 && (FULL || EVEN || ra + i - 2 < rb)) {
                        const int rout = ra + i - 2;
                        const bool interior = EVEN || (rout >= r0 && rout < r1 && col >= geo.c0 && col < geo.c1);
                        const int q = (rout - ab_lo) * Wp + kPadL + lc;
                        bwd_abv_row<NC, BLEND, NHWC, DIVM>(a, D[(i - 2) % 3], D[(i - 1) % 3], D[i % 3], V[(i - 2) % 3], V[i % 3],
                                                           rout * W + col, interior, col == 0, col + 4 == W, base_b, gin_b, gbase_b,
                                                           hw, want_gx, A + q, Bv + q);
                    }
                }
            };
            if (EE_FWD_CHUNK_IS_FULL(ra, rb, H)) body(full_t{}); else body(part_t{});
        }
    }
    if (!want_gx) return;
    __syncthreads();

    // ---- GB = fold(Sobel^T(A, Bv)) on rows [gb_lo, gb_hi) ; g_s rows [r0, r1) = fold(Gauss^T(GB)) -> g_x -------
    float* GB = R2;
    if (active) fast_stage_sobel_adjoint<R, EVEN>(geo, A, Bv, ab_lo, GB, gb_lo, gb_hi, tx, ty);
    __syncthreads();
    if (active) fast_stage_gauss_adjoint_store<NC, R, NHWC, EVEN>(a, geo, GB, gb_lo, a.e.g_x + (size_t)b * C * hw, r0, r1, tx, ty);
}

}  // namespace ee
