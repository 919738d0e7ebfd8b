// ee_edge_stream.cuh -- ROW-STREAMING kernels for the full CannyFilter / CannyFilter_BPDA (+ blend), hysteresis mode,
// for wide images (ImageNet 224 / 288 px): replaces the halo-tile kernels of ee_edge_canny_tiles.cuh, whose 64 x 64 planes
// recompute a 1.8x halo, issue 14.6 warp instructions per pixel and lose 38 % of their shared-memory wavefronts to bank
// conflicts (profiles/r1f_224_ncu_full_canny.txt).
//
// Replaces (reference): utils/core.py:222-326 (CannyFilter.forward), :426-505 (CannyFilter_BPDA.forward), the blend of the
// *_EE models (e.g. Tiny_ImageNet/models_tinyimagenet/resnet_EE.py:185-191) and the autograd graph through them.
//
// Decomposition.  One CTA owns a BAND: rows [r0, r1) of one image at FULL width (no column halo).  A thread owns one
// float4 column group for the whole band and marches down the rows; every stage of the filter is a 3-row vertical
// recurrence whose two older rows live in REGISTERS (as the horizontal partial sums the existing kernels already use), so
// no plane is ever materialised:
//
//     step t :  S(i)  ->  Bl(i-2)  ->  M, dir, gx1, gy1 (i-4)  ->  NMS + thresholds (i-6)  ->  A, Bv (i-8)  ->  GB (i-10)
//               ->  g_x (i-12)                                                     (i = first input row + t)
//
// All seven stages run in the SAME step on different rows (software pipeline, two rows of lag per stage: one for the
// 3 x 3 halo, one because a neighbour's value is visible only after the step's barrier), which leaves ONE CTA barrier per
// row.  The CTA is WARP-SPECIALISED: the "front" warps run S / blur / magnitude + direction / NMS, the "back" warps run
// hysteresis / A, Bv (forward: the blend) / the two adjoint stencils, each thread of either role owning the same column
// group (the monolithic first version needed 204 registers per thread and spilled at 128; profiles/r2a_*).  The only
// shared-memory traffic between threads of a role is the left / right neighbour of each group (two floats per plane per
// row, double buffered by step parity, conflict free); front -> back hand-over is one packed META word per group and row,
// and M / gx1 / gy1 rows in a five-row ring.  Rows are never recomputed inside a band; a band that is not the whole image
// warms the pipeline up on <= 6 halo rows per side.
//
// Input rows (x; base and g_out for the back role) are staged by the TMA engine: one elected thread issues cp.async.bulk
// (SASS UBLKCP) copies of ROW PAIRS (two consecutive rows of a channel plane are contiguous) into a ring of kStreamDepth
// pair slots, completion on an mbarrier per slot (expect_tx), plus cp.async.bulk.prefetch.L2 (UBLKPF) kStreamL2Ahead rows
// ahead; consumers read their own 16 bytes per row with one conflict-free LDS.128.  Outputs leave with 128-bit streaming
// stores (full 4 W-byte rows per channel).
//
// VAR = 0 runs CannyFilter_step125_1 (utils/core.py:549-585) through the same pipeline -- no suppression, no hysteresis: the
// NMS stage degenerates to the threshold of the row's own magnitudes, the hysteresis sums are compiled out -- for the
// BACKWARD of wide images, where the halo tiles of ee_edge_tiles.cuh reach 0.74 of the HBM peak.
//
// Arithmetic: the SAME expression trees as the other kernel families, so results are bit-identical.
#pragma once
#include "ee_edge_canny_fast.cuh"

namespace ee {

#ifndef EE_STREAM_DEPTH
#define EE_STREAM_DEPTH 2
#endif
// A second, guard-free instantiation of the step body for the steady-state rows was measured: no gain on whole-image
// bands (428 -> 435 us at 512x3x224x224) and a large loss on short bands (28-row bands 460 -> 772 us: four code regions
// of ~1000 SASS instructions each then run concurrently on one SM and thrash the instruction cache), so it is off.
#ifndef EE_STREAM_STEADY
#define EE_STREAM_STEADY 0
#endif
#ifndef EE_STREAM_L2_AHEAD
#define EE_STREAM_L2_AHEAD 6
#endif
constexpr int kStreamDepth = EE_STREAM_DEPTH;      // TMA ring slots of TWO rows each
constexpr int kStreamL2Ahead = EE_STREAM_L2_AHEAD; // L2 bulk prefetch distance in rows (0 = off)
constexpr int kStreamGRing = 5;                    // M / gx1 / gy1 ring rows (the back role reads a row four steps after the front wrote it)
constexpr int kStreamPlanes = 6;                   // neighbour-exchanged planes: S, Bl, M, A, Bv, GB

struct StreamArgs {
    FastArgs f;
    int BH;                 // rows per band
    int bands_per_img;
};

// shared-memory bytes of one CTA (host and device agree through this function)
__host__ __device__ inline size_t stream_smem_bytes(int W, int rows_per_slot, bool bwd) {
    const int GX = W / 4, GXp = (GX + 2 + 3) & ~3;
    size_t n = 128;                                                           // mbarriers
    n += (size_t)kStreamDepth * 2 * rows_per_slot * W * sizeof(float);        // TMA ring (row pairs)
    n += (size_t)kStreamPlanes * 2 * 2 * GXp * sizeof(float);                 // neighbour exchange
    n += (size_t)2 * GXp * sizeof(int);                                       // packed META rows (front -> back)
    if (!bwd) n += (size_t)2 * GX * sizeof(float4) + (size_t)2 * GXp * sizeof(int);   // M + direction rows (forward: NMS runs in the back role)
    if (bwd) n += (size_t)kStreamGRing * 3 * GX * sizeof(float4);             // M / gx1 / gy1 ring (front -> back)
    return n;
}

// ---- raw PTX: mbarrier + 1-D bulk copies (TMA engine) --------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "EE_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra EE_DONE_%=;\n"
        "bra EE_WAIT_%=;\n"
        "EE_DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar) : "memory");
}

// packed per-pixel bytes <-> ints
__device__ __forceinline__ int pack4(const int (&v)[4]) { return v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24); }

// NMS + double threshold of 4 pixels; same arithmetic as nms_threshold4 (ee_edge_canny_fast.cuh), direction words as ints
template <int VAR>
__device__ __forceinline__ void stream_nms4(const FastArgs& a, const Win& wu, const Win& wm, const Win& wd, int dirp, int (&meta)[4]) {
    constexpr bool bpda = (VAR == 2);
    float u[6], m[6], d[6];
    win_to_array(wu, u); win_to_array(wm, m); win_to_array(wd, d);
    const bool neg_lo = (0.0f > a.e.low), neg_hi = (0.0f > a.e.high);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int w = (dirp >> (8 * k)) & 15;
        const int dir = w - 1;
        const float mc = m[k + 1];
        const float mx0 = fmaxf(m[k + 2], m[k]), mx1 = fmaxf(u[k + 2], d[k]);
        const float mx2 = fmaxf(u[k + 1], d[k + 1]), mx3 = fmaxf(u[k], d[k + 2]);
        const float mx = (dir & 2) ? ((dir & 1) ? mx3 : mx2) : ((dir & 1) ? mx1 : mx0);
        const int removed = !(mc > mx);
        const float th = removed ? 0.0f : mc;
        const bool plo = th > a.e.low, phi = th > a.e.high;
        const int ilo = bpda ? (plo || (neg_lo && th <= a.e.low)) : plo;
        const int ihi = bpda ? (phi || (neg_hi && th <= a.e.high)) : phi;
        meta[k] = w | ((ilo + ihi) << 4) | (ihi << 6) | (removed << 7);
    }
}

// -------------------------------------------------------------------------------------------------------------------
// The kernel.  BWD = false: forward (x, base -> out [, edge]);  BWD = true: backward (g_out, x, base -> g_x, g_base).
// HL = halo rows of the band's first stage: 4 forward, 6 backward.  WT = image width as a compile-time constant (0 = read
// it from the arguments).  blockDim.x = 2 roles x (GX rounded up to a warp); role 0 = front, role 1 = back.
// -------------------------------------------------------------------------------------------------------------------
template <int NC, bool BLEND, int VAR, bool BWD, int WT, int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB) edge_canny_stream(const __grid_constant__ StreamArgs sa) {
    extern __shared__ __align__(128) unsigned char stream_smem[];
    constexpr int DIVM = (NC == 1) ? 0 : (NC == 3 ? 1 : 2);
    constexpr int HL = BWD ? 6 : 4;
    constexpr int NB = BLEND ? NC : 0;                    // base planes per slot
    constexpr int NG = BWD ? (BLEND ? NC : 1) : 0;        // g planes per slot
    constexpr int NROWS = NC + NB + NG;                   // planes per TMA slot (two rows each)
    const FastArgs& a = sa.f;
    const int H = a.e.H, W = WT ? WT : a.e.W, GX = W >> 2, GXp = (GX + 2 + 3) & ~3;
    // Roles alternate from warp to warp, and the alternation flips with the CTA: the SM's four schedulers take warps round
    // robin, so with "warps 0-1 front, 2-3 back" two schedulers would only ever see front warps and two only back warps,
    // and any difference in work between the roles idles half of the issue slots (profiles/r2a_*: 46 % barrier stalls).
    const int warp = (int)threadIdx.x >> 5;
    const int role = (warp ^ (int)((blockIdx.x * 2654435761u) >> 20)) & 1;      // warp-uniform; 0 = front, 1 = back
    const int tx = ((warp >> 1) << 5) + ((int)threadIdx.x & 31);
    const bool active = tx < GX;
    const int b = blockIdx.x / sa.bands_per_img;
    const int r0 = (blockIdx.x - b * sa.bands_per_img) * sa.BH, r1 = min(r0 + sa.BH, H);
    const size_t hw = (size_t)H * W;
    const int col = tx * 4;
    const bool left = (tx == 0), right = (tx == GX - 1);

    // stage k produces rows [lo_k, hi_k): k = 0 S, 1 Bl, 2 M, 3 NMS, 4 A/Bv (fwd: emit), 5 GB, 6 g_x
    auto lo_of = [&](int k) { return max(r0 - (HL - k), 0); };
    auto hi_of = [&](int k) { return min(r1 + (HL - k), H); };
    const int s_lo = lo_of(0), s_hi = hi_of(0), b_lo = lo_of(1), b_hi = hi_of(1), m_lo = lo_of(2), m_hi = hi_of(2);
    const int c_lo = lo_of(3), c_hi = hi_of(3), ab_lo = lo_of(4), ab_hi = hi_of(4);
    const int gb_lo = BWD ? lo_of(5) : 0, gb_hi = BWD ? hi_of(5) : 0;
    const int n_steps = (r1 - 1) + 2 * HL - s_lo + 1, n_pairs = (n_steps + 1) >> 1;

    // ---- shared memory carve-up
    uint64_t* bars = reinterpret_cast<uint64_t*>(stream_smem);
    float* ring = reinterpret_cast<float*>(stream_smem + 128);
    float* exch = ring + (size_t)kStreamDepth * 2 * NROWS * W;
    int* metarow = reinterpret_cast<int*>(exch + (size_t)kStreamPlanes * 4 * GXp);
    float4* gring = reinterpret_cast<float4*>(metarow + 2 * GXp);        // backward: M / gx1 / gy1 ring; forward: M rows [2][GX] + direction rows
    constexpr bool NMS_BACK = !BWD;                                      // which role runs stage 3 (balances the two roles)
    int* dirrow = reinterpret_cast<int*>(gring + 2 * GX);                // forward only
    enum { PS = 0, PBL = 1, PM = 2, PA = 3, PB = 4, PGB = 5 };
    // exchange entry of plane P, parity par: L[1 + tx] = my first value (the right neighbour of group tx - 1),
    // R[1 + tx] = my last value; L[GX + 1] / R[0] are the pad columns of the image border
    float* const ex_me = exch + 1 + tx;
    auto exL = [&](int plane, int paroff) { return ex_me + plane * 4 * GXp + paroff; };

    const float* x_b = a.e.x + (size_t)b * NC * hw;
    const float* base_b = BLEND ? a.e.base + (size_t)b * NC * hw : nullptr;
    const float* gin_b = BWD ? a.e.g_in + (size_t)b * NG * hw : nullptr;

    // ---- producer: pair slot k holds x rows s_lo + 2k, + 1 and base / g rows s_lo - 8 + 2k, + 1 (clipped to their ranges)
    auto issue_pair = [&](int k) {
        const uint32_t bar = smem_u32(&bars[k % kStreamDepth]);
        float* slot = ring + (size_t)(k % kStreamDepth) * 2 * NROWS * W;
        const uint32_t row_bytes = (uint32_t)W * sizeof(float);
        const int xa = s_lo + 2 * k, nx = max(min(xa + 2, s_hi) - xa, 0);
        const int qa = s_lo - 8 + 2 * k, ql = max(qa, ab_lo), nq = (NB + NG) > 0 ? max(min(qa + 2, ab_hi) - ql, 0) : 0;
        mbar_arrive_expect_tx(bar, (uint32_t)(NC * nx + (NB + NG) * nq) * row_bytes);
        if (nx > 0) {
#pragma unroll
            for (int c = 0; c < NC; ++c) bulk_g2s(smem_u32(slot + c * 2 * W), x_b + c * hw + (size_t)xa * W, nx * row_bytes, bar);
        }
        if (nq > 0) {
#pragma unroll
            for (int c = 0; c < NB; ++c)
                bulk_g2s(smem_u32(slot + (NC + c) * 2 * W + (ql - qa) * W), base_b + c * hw + (size_t)ql * W, nq * row_bytes, bar);
#pragma unroll
            for (int c = 0; c < NG; ++c)
                bulk_g2s(smem_u32(slot + (NC + NB + c) * 2 * W + (ql - qa) * W), gin_b + c * hw + (size_t)ql * W, nq * row_bytes, bar);
        }
        if (kStreamL2Ahead > 0) {            // the rows of pair k + kStreamL2Ahead / 2 towards L2
            const int xp = xa + kStreamL2Ahead, qp = qa + kStreamL2Ahead;
            if (xp + 1 < s_hi) {
#pragma unroll
                for (int c = 0; c < NC; ++c) l2_prefetch_bulk(x_b + c * hw + (size_t)xp * W, 2 * row_bytes);
            }
            if ((NB + NG) > 0 && qp >= ab_lo && qp + 1 < ab_hi) {
#pragma unroll
                for (int c = 0; c < NB; ++c) l2_prefetch_bulk(base_b + c * hw + (size_t)qp * W, 2 * row_bytes);
#pragma unroll
                for (int c = 0; c < NG; ++c) l2_prefetch_bulk(gin_b + c * hw + (size_t)qp * W, 2 * row_bytes);
            }
        }
    };

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kStreamDepth; ++s) mbar_init(smem_u32(&bars[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    // pad columns of the zero-extended planes (both parities) and of the META rows: written once
    if (threadIdx.x < 2) {
        const int par = threadIdx.x;
#pragma unroll
        for (int P = PM; P <= PGB; ++P) {
            float* L = exch + (size_t)(P * 4 + par * 2) * GXp;
            L[GX + 1] = 0.0f;            // right pad (read as the right neighbour of the last group)
            L[GXp] = 0.0f;               // R[0]: left pad
        }
        metarow[par * GXp] = 0;
        metarow[par * GXp + GX + 1] = 0;
    }
    __syncthreads();
    const bool producer = (role == 1 && tx == 0);
    if (producer) {
        for (int k = 0; k < kStreamDepth && k < n_pairs; ++k) issue_pair(k);
    }

    const float c0g = a.e.c0, c1g = a.e.c1, c2g = a.e.c2, fC = a.e.fC, wgt = a.e.w;
    const AdjBorder bd = {left, right};

    // ---- stage 3 (either role): non-maximum suppression + double threshold of row n = m - 1, from the M row pushed
    //      this step (own values mo, packed directions dir_in; neighbours from the exchange)
    Win w0 = zero_win(), w1 = zero_win();                               // M windows of rows m-2, m-1
    int dir_prev = 0;                                                   // packed direction words of row m-1
    float m_prev[4] = {0.0f, 0.0f, 0.0f, 0.0f};                         // VAR 0: own magnitudes of row m-1
    auto stage3 = [&](auto steady_tag, const int t, const int prv, const float (&mo)[4], const int dir_in) {
        constexpr bool ST = decltype(steady_tag)::value;
        const int m = s_lo + t - 5;                                     // M row pushed (produced by stage 2 last step)
        if (ST || (m >= m_lo && m <= m_hi)) {
            if constexpr (VAR == 0) {
                // CannyFilter_step125_1: edge = To_compare(magm, high) of the pixel itself (core.py:578-583); META carries
                // the bit (and a NaN flag: To_compare leaves NaN alone) with the same one-row lag as the Canny variants
                const int n = m - 1;
                if (ST || (n >= c_lo && n < c_hi)) {
                    const bool neg_hi = (0.0f > a.e.high);
                    int meta[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float v = m_prev[k];
                        const int ihi = (v > a.e.high) || (neg_hi && v <= a.e.high);
                        meta[k] = ((v != v) ? 1 : 0) | ((2 * ihi) << 4) | (ihi << 6);
                    }
                    metarow[(t & 1) * GXp + 1 + tx] = pack4(meta);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) m_prev[k] = (ST || m < m_hi) ? mo[k] : 0.0f;
            } else {
                Win w = zero_win();
                if (ST || m < m_hi) {
                    const float* L = exL(PM, prv);
                    w.l = L[GXp - 1]; w.r = L[1];
                    w.m0 = mo[0]; w.m1 = mo[1]; w.m2 = mo[2]; w.m3 = mo[3];
                }
                const int n = m - 1;
                if (ST || (n >= c_lo && n < c_hi)) {
                    int meta[4];
                    stream_nms4<VAR>(a, w0, w1, w, dir_prev, meta);
                    metarow[(t & 1) * GXp + 1 + tx] = pack4(meta);
                }
                w0 = w1; w1 = w;
                dir_prev = dir_in;
            }
        }
    };

    if (role == 0) {
        // ================================================ FRONT ROLE ================================================
        float s_own[4] = {0, 0, 0, 0};                                  // S row pushed into the blur next step
        float bT[4] = {0, 0, 0, 0}, bP[4] = {0, 0, 0, 0};               // blur: T = P(i-1) + Q(i), P(i)
        float bl_own[4] = {0, 0, 0, 0};
        float D0[4], D1[4], V0[4], V1[4];                               // Sobel partials of Bl rows j-1, j
#pragma unroll
        for (int k = 0; k < 4; ++k) { D0[k] = D1[k] = V0[k] = V1[k] = 0.0f; }
        float m_own[4] = {0, 0, 0, 0};                                  // M row pushed into the NMS next step
        int dir_new = 0;                                                // packed direction words of row m
        int wslot = 0;                                                  // ring slot of the next M / gx1 / gy1 row
        // steady state: every stage has a row in range and none of them touches the first / last row of its range, so
        // all the range tests below are compile-time true (ST) and the stage bodies become one branch-free block that
        // the scheduler can interleave; the warm-up / drain steps run the guarded instantiation
        const int t_st0 = NMS_BACK ? max(max(2, b_lo - s_lo + 4), m_lo - s_lo + 4)
                                   : max(max(2, b_lo - s_lo + 4), max(m_lo - s_lo + 5, c_lo - s_lo + 6));
        const int t_st1 = NMS_BACK ? min(min(s_hi - s_lo, b_hi - s_lo + 2), m_hi - s_lo + 4)
                                   : min(min(s_hi - s_lo, b_hi - s_lo + 2), min(m_hi - s_lo + 4, c_hi - s_lo + 6));
        auto front_step = [&](auto steady_tag, const int t) {
            constexpr bool ST = decltype(steady_tag)::value;
            const int cur = (t & 1) * 2 * GXp, prv = 2 * GXp - cur;
            if ((t & 1) == 0) mbar_wait(smem_u32(&bars[(t >> 1) % kStreamDepth]), (uint32_t)(((t >> 1) / kStreamDepth) & 1));
            if (active) {
                // ---- stage 3: non-maximum suppression + double threshold of row n = m - 1
                if constexpr (!NMS_BACK) stage3(steady_tag, t, prv, m_own, dir_new);
                // ---- stage 2: Sobel, magnitude, direction of row m = j - 1
                {
                    const int j = s_lo + t - 3;                    // Bl row pushed (produced by the blur last step)
                    if (ST || (j >= b_lo && j <= b_hi)) {
                        float D[4], V[4];
                        if (ST || j < b_hi) {
                            const float* L = exL(PBL, prv);
                            Win w; w.l = L[GXp - 1]; w.r = L[1];
                            w.m0 = bl_own[0]; w.m1 = bl_own[1]; w.m2 = bl_own[2]; w.m3 = bl_own[3];
                            sobel_partials(w, D, V);
                        } else {                                    // replicate below the image: row H := row H-1
#pragma unroll
                            for (int k = 0; k < 4; ++k) { D[k] = D1[k]; V[k] = V1[k]; }
                        }
                        if (!ST && j == b_lo) {                    // first row: replicate above (row -1 := row 0 when b_lo == 0)
#pragma unroll
                            for (int k = 0; k < 4; ++k) { D1[k] = D[k]; V1[k] = V[k]; }
                        } else {
                            const int m = j - 1;
                            if (ST || (m >= m_lo && m < m_hi)) {
                                float sgx[4], sgy[4], gx1[4], gy1[4], mm[4];
                                int dw[4];
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    sgx[k] = fmaf(0.5f, D0[k] + D[k], D1[k]);
                                    sgy[k] = V[k] - V0[k];
                                }
                                div_channels8<DIVM>(sgx, sgy, fC, gx1, gy1);
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const float mag = magnitude(gx1[k], gy1[k]);
                                    mm[k] = (VAR != 2 && mag < a.e.alpha) ? 0.0f : mag;
                                    dw[k] = (VAR == 0) ? 0 : orient_dir(gx1[k], gy1[k]) + 1;
                                    m_own[k] = mm[k];
                                }
                                dir_new = pack4(dw);
                                float* L = exL(PM, cur);
                                L[0] = mm[0]; L[GXp] = mm[3];
                                if constexpr (NMS_BACK) {
                                    gring[(t & 1) * GX + tx] = make_float4(mm[0], mm[1], mm[2], mm[3]);
                                    dirrow[(t & 1) * GXp + tx] = dir_new;
                                }
                                if constexpr (BWD) {
                                    float4* gr = gring + (size_t)wslot * 3 * GX + tx;
                                    gr[0] = make_float4(mm[0], mm[1], mm[2], mm[3]);
                                    gr[GX] = make_float4(gx1[0], gx1[1], gx1[2], gx1[3]);
                                    gr[2 * GX] = make_float4(gy1[0], gy1[1], gy1[2], gy1[3]);
                                    wslot = (wslot == kStreamGRing - 1) ? 0 : wslot + 1;
                                }
                            }
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) { D0[k] = D1[k]; D1[k] = D[k]; V0[k] = V1[k]; V1[k] = V[k]; }
                    }
                }
                // ---- stage 1: 3 x 3 Gaussian blur of row j = i - 1
                {
                    const int i = s_lo + t - 1;                    // S row pushed (produced by stage 0 last step)
                    if (ST || (t >= 1 && i <= s_hi)) {
                        float P[4], Q[4];
                        if (ST || i < s_hi) {
                            const float* L = exL(PS, prv);
                            Win w; w.l = L[GXp - 1]; w.r = L[1];
                            w.m0 = s_own[0]; w.m1 = s_own[1]; w.m2 = s_own[2]; w.m3 = s_own[3];
                            gauss_partials(w, c0g, c1g, c2g, P, Q);
                        } else {                                    // replicate below the image
#pragma unroll
                            for (int k = 0; k < 4; ++k) { P[k] = bP[k]; Q[k] = bP[k]; }
                        }
                        if (!ST && i == s_lo) {                    // first row: P(i-1) := P(i)
#pragma unroll
                            for (int k = 0; k < 4; ++k) { bT[k] = P[k] + Q[k]; bP[k] = P[k]; }
                        } else {
                            const int j = i - 1;
                            if (ST || (j >= b_lo && j < b_hi)) {
#pragma unroll
                                for (int k = 0; k < 4; ++k) bl_own[k] = bT[k] + P[k];         // (P(j-1) + Q(j)) + P(j+1)
                                float* L = exL(PBL, cur);
                                L[0] = bl_own[0]; L[GXp] = bl_own[3];
                                if (left) L[GXp - 1] = bl_own[0];
                                if (right) L[1] = bl_own[3];
                            }
#pragma unroll
                            for (int k = 0; k < 4; ++k) { bT[k] = bP[k] + Q[k]; bP[k] = P[k]; }
                        }
                    }
                }
                // ---- stage 0: channel sum of the x row the TMA ring delivered
                {
                    const int i0 = s_lo + t;
                    if (ST || i0 < s_hi) {
                        const float* slot = ring + (size_t)((t >> 1) % kStreamDepth) * 2 * NROWS * W + (t & 1) * W + col;
                        float4 acc = *reinterpret_cast<const float4*>(slot);
#pragma unroll
                        for (int c = 1; c < NC; ++c) acc = f4add(acc, *reinterpret_cast<const float4*>(slot + c * 2 * W));
                        s_own[0] = acc.x; s_own[1] = acc.y; s_own[2] = acc.z; s_own[3] = acc.w;
                        float* L = exL(PS, cur);
                        L[0] = acc.x; L[GXp] = acc.w;
                        if (left) L[GXp - 1] = acc.x;
                        if (right) L[1] = acc.w;
                    }
                }
            }
            asm volatile("bar.sync 0;" ::: "memory");
        };
        for (int t = 0; t < n_steps; ++t) {
            if (EE_STREAM_STEADY && t >= t_st0 && t < t_st1) front_step(full_t{}, t); else front_step(part_t{}, t);
        }
    } else {
        // ================================================ BACK ROLE =================================================
        int hs0 = 0, hs1 = 0, cw1 = 0;                                  // hysteresis: packed 3-sums of rows q-1, q ; META of row q
        int rslot = (ab_lo - m_lo) % kStreamGRing;                      // ring slot of the next A / Bv row
        float a_own[4] = {0, 0, 0, 0}, b_own[4] = {0, 0, 0, 0};
        float HA0[4], HA1[4], HB0[4], HB1[4], HAr0 = 0, HAr1 = 0, HBr0 = 0, HBr1 = 0;
        float gb_own[4] = {0, 0, 0, 0};
        float gP0[4], gP1[4], gQ1[4], gPr0 = 0, gPr1 = 0, gQr1 = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) { HA0[k] = HA1[k] = HB0[k] = HB1[k] = 0.0f; gP0[k] = gP1[k] = gQ1[k] = 0.0f; }
        (void)rslot; (void)a_own; (void)b_own; (void)gb_own; (void)HAr0; (void)HAr1; (void)HBr0; (void)HBr1; (void)gPr0; (void)gPr1; (void)gQr1;
        // steady state of the back role (see the front role): rows strictly inside the image and inside the band
        const int t_st0 = BWD ? max(max(c_lo - s_lo + 7, ab_lo - s_lo + 9), max(max(gb_lo, 1) - s_lo + 11, max(r0, 1) - s_lo + 12))
                              : max(max(c_lo - s_lo + 7, ab_lo - s_lo + 8), max(m_lo - s_lo + 5, c_lo - s_lo + 6));
        const int t_st1 = BWD ? min(min(c_hi - s_lo + 7, ab_hi - s_lo + 8), min(min(gb_hi, H - 1) - s_lo + 10, min(r1, H - 1) - s_lo + 12))
                              : min(min(c_hi - s_lo + 7, ab_hi - s_lo + 8), min(m_hi - s_lo + 5, c_hi - s_lo + 6));
        auto back_step = [&](auto steady_tag, const int t) {
            constexpr bool ST = decltype(steady_tag)::value;
            const int cur = (t & 1) * 2 * GXp, prv = 2 * GXp - cur;
            if ((t & 1) == 0) {
                const int k = t >> 1;
                if (producer && k >= 1 && k + kStreamDepth - 1 < n_pairs) issue_pair(k + kStreamDepth - 1);
                mbar_wait(smem_u32(&bars[k % kStreamDepth]), (uint32_t)((k / kStreamDepth) & 1));
            }
            if (active) {
                // ---- stage 6 (backward): g_x row p = fold(Gauss^T(GB))
                if constexpr (BWD) {
                    const int pp = s_lo + t - 11;                      // GB row pushed (produced by stage 5 last step)
                    if (a.e.g_x != nullptr && (ST || (pp >= gb_lo && pp <= gb_hi))) {
                        float P[4] = {0, 0, 0, 0}, Q[4] = {0, 0, 0, 0}, Pr = 0.0f, Qr = 0.0f;
                        if (ST || pp < gb_hi) {
                            const float* L = exL(PGB, prv);
                            Win w; w.l = L[GXp - 1]; w.r = L[1];
                            w.m0 = gb_own[0]; w.m1 = gb_own[1]; w.m2 = gb_own[2]; w.m3 = gb_own[3];
                            gauss_adj_partials(w, bd, c0g, c1g, c2g, P, Q, Pr, Qr);
                        }
                        const int p = pp - 1;
                        if (ST || (p >= r0 && p < r1)) {
                            float o[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) o[k] = (gP0[k] + gQ1[k]) + P[k];
                            if (left || right) {
                                const float tt = (gPr0 + gQr1) + Pr;
                                if (left) o[0] = o[0] + tt; else o[3] = o[3] + tt;
                            }
                            if (!ST && p == 0) {                       // ring row -1 = (zero, zero, row 0), folded into row 0
                                float h[4];
#pragma unroll
                                for (int k = 0; k < 4; ++k) h[k] = (0.0f + 0.0f) + gP1[k];
                                if (left || right) {
                                    const float tt = (0.0f + 0.0f) + gPr1;
                                    if (left) h[0] = h[0] + tt; else h[3] = h[3] + tt;
                                }
#pragma unroll
                                for (int k = 0; k < 4; ++k) o[k] = o[k] + h[k];
                            }
                            if (!ST && p == H - 1) {                   // ring row H = (row H-1, zero, zero), folded into row H-1
                                float h[4];
#pragma unroll
                                for (int k = 0; k < 4; ++k) h[k] = (gP1[k] + 0.0f) + 0.0f;
                                if (left || right) {
                                    const float tt = (gPr1 + 0.0f) + 0.0f;
                                    if (left) h[0] = h[0] + tt; else h[3] = h[3] + tt;
                                }
#pragma unroll
                                for (int k = 0; k < 4; ++k) o[k] = o[k] + h[k];
                            }
                            const float4 v = make_float4(o[0], o[1], o[2], o[3]);
                            float* pg = a.e.g_x + (size_t)b * NC * hw + (size_t)p * W + col;
#pragma unroll
                            for (int c = 0; c < NC; ++c) __stcs(reinterpret_cast<float4*>(pg + c * hw), v);
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) { gP0[k] = gP1[k]; gP1[k] = P[k]; gQ1[k] = Q[k]; }
                        gPr0 = gPr1; gPr1 = Pr; gQr1 = Qr;
                    }
                }
                // ---- stage 5 (backward): GB row p = fold(Sobel^T(A, Bv))
                if constexpr (BWD) {
                    const int q = s_lo + t - 9;                        // A / Bv row pushed
                    if (a.e.g_x != nullptr && (ST || (q >= ab_lo && q <= ab_hi))) {
                        float HA[4] = {0, 0, 0, 0}, HB[4] = {0, 0, 0, 0}, HAr = 0.0f, HBr = 0.0f;
                        if (ST || q < ab_hi) {
                            const float* LA = exL(PA, prv);
                            const float* LB = exL(PB, prv);
                            Win wa, wb;
                            wa.l = LA[GXp - 1]; wa.r = LA[1]; wb.l = LB[GXp - 1]; wb.r = LB[1];
                            wa.m0 = a_own[0]; wa.m1 = a_own[1]; wa.m2 = a_own[2]; wa.m3 = a_own[3];
                            wb.m0 = b_own[0]; wb.m1 = b_own[1]; wb.m2 = b_own[2]; wb.m3 = b_own[3];
                            sobel_adj_partials(wa, wb, bd, HA, HB, HAr, HBr);
                        }
                        const int p = q - 1;
                        if (ST || (p >= gb_lo && p < gb_hi)) {
                            float o[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) o[k] = fmaf(0.5f, HA0[k] + HA[k], HA1[k]) + (HB0[k] - HB[k]);
                            if (left || right) {
                                const float tt = fmaf(0.5f, HAr0 + HAr, HAr1) + (HBr0 - HBr);
                                if (left) o[0] = o[0] + tt; else o[3] = o[3] + tt;
                            }
                            if (!ST && p == 0) {                       // ring row -1 = (zero, zero, row 0)
                                float h[4];
#pragma unroll
                                for (int k = 0; k < 4; ++k) h[k] = fmaf(0.5f, 0.0f + HA1[k], 0.0f) + (0.0f - HB1[k]);
                                if (left || right) {
                                    const float tt = fmaf(0.5f, 0.0f + HAr1, 0.0f) + (0.0f - HBr1);
                                    if (left) h[0] = h[0] + tt; else h[3] = h[3] + tt;
                                }
#pragma unroll
                                for (int k = 0; k < 4; ++k) o[k] = o[k] + h[k];
                            }
                            if (!ST && p == H - 1) {                   // ring row H = (row H-1, zero, zero)
                                float h[4];
#pragma unroll
                                for (int k = 0; k < 4; ++k) h[k] = fmaf(0.5f, HA1[k] + 0.0f, 0.0f) + (HB1[k] - 0.0f);
                                if (left || right) {
                                    const float tt = fmaf(0.5f, HAr1 + 0.0f, 0.0f) + (HBr1 - 0.0f);
                                    if (left) h[0] = h[0] + tt; else h[3] = h[3] + tt;
                                }
#pragma unroll
                                for (int k = 0; k < 4; ++k) o[k] = o[k] + h[k];
                            }
#pragma unroll
                            for (int k = 0; k < 4; ++k) gb_own[k] = o[k];
                            float* L = exL(PGB, cur);
                            L[0] = o[0]; L[GXp] = o[3];
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) { HA0[k] = HA1[k]; HA1[k] = HA[k]; HB0[k] = HB1[k]; HB1[k] = HB[k]; }
                        HAr0 = HAr1; HAr1 = HAr; HBr0 = HBr1; HBr1 = HBr;
                    }
                }
                // ---- stage 4: hysteresis of row q = np - 1; forward: blend and store, backward: A / Bv
                {
                    const int np = s_lo + t - 7;                       // META row pushed (produced by the NMS last step)
                    if (ST || (np >= c_lo && np <= c_hi)) {
                        int hs = 0, cwn = 0;
                        if (ST || np < c_hi) {
                            const int* mr = metarow + ((t & 1) ^ 1) * GXp + 1 + tx;
                            cwn = mr[0];
                            if constexpr (VAR != 0) {
                                const int lhp = (cwn >> 4) & 0x03030303;
                                hs = lhp + ((lhp << 8) | ((mr[-1] >> 28) & 3)) + ((lhp >> 8) | (((mr[1] >> 4) & 3) << 24));
                            }
                        }
                        const int q = np - 1;
                        if (ST || (q >= ab_lo && q < ab_hi)) {
                            const int nsum = hs0 + hs1 + hs;
                            int meta[4], wih[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                meta[k] = (cw1 >> (8 * k)) & 255;
                                wih[k] = (VAR != 0) && (meta_lh(meta[k]) == 1) && (((nsum >> (8 * k)) & 255) >= 2);
                            }
                            const int pix = q * W + col;
                            const float* slot = ring + (size_t)((t >> 1) % kStreamDepth) * 2 * NROWS * W + (t & 1) * W + col;
                            if constexpr (!BWD) {
                                float e[4];
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    e[k] = (VAR == 0 && (meta[k] & 1)) ? __int_as_float(0x7fc00000) : (float)(meta_hi(meta[k]) + wih[k]);
                                if (a.e.edge) __stcs(reinterpret_cast<float4*>(a.e.edge + (size_t)b * hw + pix), make_float4(e[0], e[1], e[2], e[3]));
                                if constexpr (BLEND) {
                                    const float w0_ = wgt * e[0], w1_ = wgt * e[1], w2_ = wgt * e[2], w3_ = wgt * e[3];
                                    float* out_b = a.e.out + (size_t)b * NC * hw + pix;
#pragma unroll
                                    for (int c = 0; c < NC; ++c) {
                                        const float4 bs = *reinterpret_cast<const float4*>(slot + (NC + c) * 2 * W);
                                        __stcs(reinterpret_cast<float4*>(out_b + c * hw),
                                               make_float4(clamp01_fast(bs.x + w0_), clamp01_fast(bs.y + w1_), clamp01_fast(bs.z + w2_),
                                                           clamp01_fast(bs.w + w3_)));
                                    }
                                }
                            } else {
                                const float4* gr = gring + (size_t)rslot * 3 * GX + tx;
                                const float4 tm4 = gr[0], tx4 = gr[GX], ty4 = gr[2 * GX];
                                rslot = (rslot == kStreamGRing - 1) ? 0 : rslot + 1;
                                const float mag[4] = {tm4.x, tm4.y, tm4.z, tm4.w};
                                const float gx1[4] = {tx4.x, tx4.y, tx4.z, tx4.w}, gy1[4] = {ty4.x, ty4.y, ty4.z, ty4.w};
                                float thin[4], ge[4];
#pragma unroll
                                for (int k = 0; k < 4; ++k) thin[k] = meta_removed(meta[k]) ? 0.0f : mag[k];
                                if constexpr (BLEND) {
                                    float we[4];
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                        we[k] = wgt * ((VAR == 0 && (meta[k] & 1)) ? __int_as_float(0x7fc00000) : (float)(meta_hi(meta[k]) + wih[k]));
                                    const bool interior = (q >= r0 && q < r1);
                                    float* gbase_p = a.e.g_base ? a.e.g_base + (size_t)b * NC * hw + pix : nullptr;
#pragma unroll
                                    for (int c = 0; c < NC; ++c) {
                                        const float4 bsc = *reinterpret_cast<const float4*>(slot + (NC + c) * 2 * W);
                                        const float4 goc = *reinterpret_cast<const float4*>(slot + (NC + NB + c) * 2 * W);
                                        const float bsv[4] = {bsc.x, bsc.y, bsc.z, bsc.w}, gov[4] = {goc.x, goc.y, goc.z, goc.w};
                                        float gp[4];
#pragma unroll
                                        for (int k = 0; k < 4; ++k) {
                                            const float pre = bsv[k] + we[k];
                                            gp[k] = (pre >= 0.0f && pre <= 1.0f) ? gov[k] : 0.0f;
                                            ge[k] = (c == 0) ? gp[k] * wgt : fmaf(gp[k], wgt, ge[k]);
                                        }
                                        if (gbase_p && interior) __stcs(reinterpret_cast<float4*>(gbase_p + c * hw), make_float4(gp[0], gp[1], gp[2], gp[3]));
                                    }
                                } else {
                                    const float4 tg = *reinterpret_cast<const float4*>(slot + NC * 2 * W);
                                    ge[0] = tg.x; ge[1] = tg.y; ge[2] = tg.z; ge[3] = tg.w;
                                }
                                // mag_backward (ee_device.cuh) for the 4 pixels at once: ONE branch per group instead of four
                                // divergent ones (NMS keeps ~25 % of the pixels, so every warp used to run all four IEEE
                                // divisions one after the other); inside, the four divisions are independent and branch-free
                                float av[4] = {0.0f, 0.0f, 0.0f, 0.0f}, bv[4] = {0.0f, 0.0f, 0.0f, 0.0f}, gm[4];
                                bool any = false;
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    gm[k] = g_thin_of_v(VAR, a.e.low, a.e.high, MODE_HYST, ge[k], thin[k], wih[k]);
                                    if (meta_removed(meta[k]) || mag[k] == 0.0f) gm[k] = 0.0f;
                                    any = any || (gm[k] != 0.0f);
                                }
                                if (any) {
#pragma unroll
                                    for (int k = 0; k < 4; ++k) {
                                        const bool on = (gm[k] != 0.0f);
                                        const float tq = gm[k] / (on ? mag[k] * fC : 1.0f);
                                        av[k] = on ? tq * gx1[k] : 0.0f;
                                        bv[k] = on ? tq * gy1[k] : 0.0f;
                                    }
                                }
#pragma unroll
                                for (int k = 0; k < 4; ++k) { a_own[k] = av[k]; b_own[k] = bv[k]; }
                                float* LA = exL(PA, cur);
                                float* LB = exL(PB, cur);
                                LA[0] = av[0]; LA[GXp] = av[3];
                                LB[0] = bv[0]; LB[GXp] = bv[3];
                            }
                        }
                        hs0 = hs1; hs1 = hs; cw1 = cwn;
                    }
                }
                // ---- stage 3 (forward: NMS runs here so that the two roles carry similar work)
                if constexpr (NMS_BACK) {
                    const int m = s_lo + t - 5;
                    float mo[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                    int din = 0;
                    if (ST || (m >= m_lo && m < m_hi)) {
                        const float4 v = gring[((t & 1) ^ 1) * GX + tx];
                        mo[0] = v.x; mo[1] = v.y; mo[2] = v.z; mo[3] = v.w;
                        din = dirrow[((t & 1) ^ 1) * GXp + tx];
                    }
                    stage3(steady_tag, t, prv, mo, din);
                }
            }
            asm volatile("bar.sync 0;" ::: "memory");
        };
        for (int t = 0; t < n_steps; ++t) {
            if (EE_STREAM_STEADY && t >= t_st0 && t < t_st1) back_step(full_t{}, t); else back_step(part_t{}, t);
        }
    }
}

}  // namespace ee
