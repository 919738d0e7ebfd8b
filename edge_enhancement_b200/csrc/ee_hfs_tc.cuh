// ee_hfs_tc.cuh -- HighFreqSuppress (utils/core.py:15-55) for 64 x 64 planes, radius 8 (Tiny-ImageNet), with the four
// dense products of ee_hfs.cuh on the 5th-generation tensor cores:
//
//     T = x CB        D^T = T^T RB        G = W o D (+ the four cross terms of frequency row -r)        V = RB G        y = V CB^T
//
// `tcgen05.mma.cta_group::1.kind::tf32`, operands in shared memory (K-major, no swizzle: 8 x 16 B core matrices),
// accumulators in tensor memory, read back with `tcgen05.ld`.  fp32 accuracy comes from the 3 x TF32 split a = a_hi + a_lo
// (a_hi = a rounded to 11 significant bits, a - a_hi exact): a b ~ a_hi b_hi + a_lo b_hi + a_hi b_lo, accumulated in fp32 by
// the tensor core.  Unlike the FFMA kernel this one is NOT bit-identical to the C oracle (the accumulation order inside the
// tensor core is not specified): measured 1.3e-6 max abs error against float64 on [0,1] inputs (FFMA kernel: 0.6e-6);
// tests/test_gpu_hfs.py bounds it.  Entry point: ee_hfs_tc_f32 (opt-in; ee_hfs_f32 keeps the bit-exact FFMA kernel).
//
// One CTA of 256 threads works on a PAIR of planes at a time, persistent over pairs, two CTAs per SM (108 KB of shared
// memory, 256 of the 512 tensor-memory columns each).  Warps w and w + 4 share tensor-memory lanes 32 (w % 4) .. + 31 and
// take half of the columns each.
//
//   0. x of the pair arrives by TMA: two tensor copies (32 floats x 128 rows, 128 B swizzle) write it straight into the
//      K-major SWIZZLE_128B operand layout of the lo operand's buffer, issued one iteration ahead (as soon as product 1 of
//      the previous pair has been committed); the split runs in place: hi to work area A, lo back to the same offset.
//   1. T[128 x 16] = X CB                       M = 128, N = 32 | 16, K = 64    (B = [CB_hi ; CB_lo] stacked along N: the
//                                               reader adds column 16 + j to column j; 2 MMAs per K step instead of 3)
//   2. each thread reads its row of T from tensor memory and stores it TRANSPOSED: rows m = (plane, j), K = h.
//   3. D^T[(plane, j) x 24] = T^T RB            M = 64 (32 used), N = 48 | 24, K = 64
//   4. 16 lanes per plane read D, form G (elementwise weight + the cross terms, two shuffles), store rows n = (plane, j), K = i.
//   5. V[64 x (plane, j)] = RB G                M = 64, N = 32, K = 24           (3 MMAs per K step)
//   6. V -> rows m = (plane, h), K = j (16-lane tensor-memory loads: the M = 64 accumulator uses lanes 0..15 of a quadrant).
//   7. y[128 x 64] = V CB^T                     M = 128, N = 64, K = 16          (3 MMAs per K step)
//   8. output: tensor memory -> the 128 B-swizzled layout in work area A -> one TMA tensor store of 32 x 32 floats per warp
//      (a TMA REDUCTION store, y += tile, when `add` aliases y: the in-place accumulation of the front end's backward);
//      with a separate `add` buffer: 272 B-strided rows -> + add -> coalesced 128-bit stores.
//
// The constant operands (CB, RB in both orientations, hi and lo: 40 KB) are built once per CTA from the caller's tables.
// T~, G~, V~ and the output staging share work area A (each dead before the next is written).  Each product is issued by
// thread 0 (descriptors precomputed, a K step adds a constant to the address field) and committed to one mbarrier that all
// threads wait on; generic-proxy writes are fenced (fence.proxy.async) before the barrier that precedes the issue.
//
// Measured (profiles/README.md, r2z): 103 us at 4096x3x64x64 against 153 us for the FFMA kernel (3.0x fewer warp
// instructions).  Global memory is touched by the TMA engine only: with LDG / STG for x and y the same kernel took 151 us --
// the bursts of global requests sat in front of the shared- and tensor-memory traffic of the other stages in the load /
// store unit (ablation in profiles/README.md).
#pragma once
#include "ee_edge_stream.cuh"      // smem_u32, mbarrier helpers
#include "ee_hfs.cuh"
#ifdef EE_TC_PROFILE
#include <cstdio>
#endif

namespace ee {
namespace hfs_tc {

constexpr int N = 64, R = 8, NJ = 2 * R - 1, NI = 2 * R + 1;
constexpr int NJp = 16, NIp = 20;            // strides of the caller's tables (HfsDims<64, 8>)
constexpr int NJt = 16, NIt = 24;            // padded to the MMA shapes
constexpr int kThreads = 256;
constexpr int kTmemCols = 256;
constexpr uint32_t COL_T = 0, COL_D = 32, COL_V = 128, COL_Y = 0;      // widths 32, 48 (hi and lo halves), 32, 64; y reuses T / D

// operand images: byte offset of element (row, k) = (row / 8) * SBO + (k / 4) * LBO + (row % 8) * 16 + (k % 4) * 4
constexpr uint32_t X_LBO = 128, X_SBO = 16 * 128, X_BYTES = 16 * X_SBO;            // 128 rows x 64
constexpr uint32_t TT_LBO = 144, TT_SBO = 16 * 144;                                // 64 rows x 64 (odd chunk stride: the
                                                                                   // transposing STS.32 are conflict-free)
constexpr uint32_t CBT_LBO = 128, CBT_SBO = 16 * 128, CBT_BYTES = 2 * CBT_SBO;     // 16 rows (j) x 64 (w)
constexpr uint32_t RBT_LBO = 128, RBT_SBO = 16 * 128, RBT_BYTES = 3 * RBT_SBO;     // 24 rows (i) x 64 (h)
constexpr uint32_t RB_LBO = 128, RB_SBO = 6 * 128, RB_BYTES = 8 * RB_SBO;          // 64 rows (h) x 24 (i)
constexpr uint32_t G_LBO = 128, G_SBO = 6 * 128, G_BYTES = 4 * G_SBO;              // 32 rows (plane, j) x 24 (i)
constexpr uint32_t V_LBO = 128, V_SBO = 4 * 128;                                   // 128 rows (plane, h) x 16 (j)
constexpr uint32_t CB_LBO = 128, CB_SBO = 4 * 128, CB_BYTES = 8 * CB_SBO;          // 64 rows (w) x 16 (j)
constexpr uint32_t Y_ROW = 272;                     // y staging rows: 68 floats, 4 mod 32 words
static_assert(128 * Y_ROW <= X_BYTES + 2048, "T~ / V~ alias the x operand");

constexpr uint32_t OFF_W = 128;
constexpr uint32_t OFF_CBT = OFF_W + NIt * NJt * 4;         // hi, then lo
constexpr uint32_t OFF_RBT = OFF_CBT + 2 * CBT_BYTES;
constexpr uint32_t OFF_RB = OFF_RBT + 2 * RBT_BYTES;
constexpr uint32_t OFF_CB = OFF_RB + 2 * RB_BYTES;
// work area A (x_hi operand; then T~ hi | lo, G~ hi | lo, V~ hi | lo, the y staging rows -- each dead before the next is
// written), 2 KB of spill for the 272 B-strided y rows, work area B (x_lo operand ONLY: free from the commit of product 1 to
// the next pair, which is when the next pair's x is copied into it asynchronously)
constexpr uint32_t OFF_XH = (OFF_CB + 2 * CB_BYTES + 1023u) & ~1023u;      // SWIZZLE_128B operands: 1024 B aligned
constexpr uint32_t OFF_XL = OFF_XH + X_BYTES + 2048;
constexpr uint32_t kSmem = OFF_XL + X_BYTES;                // 110592 B: two CTAs per SM
static_assert(OFF_XH % 1024 == 0 && OFF_XL % 1024 == 0, "swizzle atoms");
constexpr uint32_t XSW_KB = 128 * 128;                      // one K block (32 floats = 128 B per row) of the swizzled x operand
constexpr uint32_t A_TT_LO = 4 * TT_SBO;                    // T~ lo behind the 32 valid rows of T~ hi
constexpr uint32_t A_G_HI = 20480, A_G_LO = A_G_HI + G_BYTES;
constexpr uint32_t A_V_LO = 16 * V_SBO;
static_assert(A_TT_LO + 8 * TT_SBO <= X_BYTES, "the M = 64 read of T~ lo stays inside work area A");
static_assert(A_G_LO + G_BYTES <= X_BYTES && A_G_HI >= 2 * A_TT_LO && A_G_HI >= 2 * A_V_LO, "G~ clear of the valid T~ / V~ rows");

__device__ __forceinline__ uint32_t op_off(int row, int k, uint32_t lbo, uint32_t sbo) {
    return (uint32_t)(row >> 3) * sbo + (uint32_t)(k >> 2) * lbo + (uint32_t)(row & 7) * 16u + (uint32_t)(k & 3) * 4u;
}
// shared-memory matrix descriptor, K-major, SWIZZLE_NONE: start address, leading (K chunk) and stride (8-row group) byte
// offsets in 16 B units, descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// K-major SWIZZLE_128B: rows of 128 B, 16 B chunks XOR-ed with (row % 8), 8-row groups 1024 B apart (what a TMA tensor copy
// with CU_TENSOR_MAP_SWIZZLE_128B writes); leading offset unused (1), layout type 2
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D = f32 (bit 4), A = B = tf32 (2 at bits 7 and 10), both K-major, N / 8 at bit 17, M / 16 at bit 24
__host__ __device__ constexpr uint32_t idesc(int M, int Nn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(Nn >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// Veltkamp split at 13 bits: hi = v rounded to 11 significant bits (a TF32 number), lo = v - hi exactly (|lo| <= 2^-12 |v|);
// the tensor core reads only the TF32 part of lo (error <= 2^-23 |v|).  Four FP32 instructions, no cvt.
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
    const float c = v * 8193.0f;
    hi = c - (c - v);
    lo = v - hi;
}
template <bool ACC>
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t id) {
    if (ACC)
        asm volatile("{\n.reg .pred p;\nsetp.eq.u32 p, 1, 1;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
                     "l"(adesc), "l"(bdesc), "r"(id) : "memory");
    else
        asm volatile("{\n.reg .pred p;\nsetp.eq.u32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
                     "l"(adesc), "l"(bdesc), "r"(id) : "memory");
}
// D[tmem][:, 0 .. 2 NB) = A_hi [B_hi ; B_lo]^T and D[:, 0 .. NB) += A_lo B_hi^T over KS K steps of 8 elements (two 16 B chunks):
// the hi and lo images of B are contiguous, so one MMA of width 2 NB reads A_hi once for both; the reader adds column
// NB + n to column n (3 x TF32 with two MMAs per K step instead of three).  Descriptors are built once per kernel; a K step
// adds a compile-time constant to their address field (the issuing thread is alone: every instruction it spends is latency).
template <int KS, uint32_t A_LBO, uint32_t B_LBO>
__device__ __forceinline__ void mma_split(uint32_t d_tmem, uint64_t a_hi, uint64_t a_lo, uint64_t b, uint32_t id_wide, uint32_t id_narrow) {
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
        const uint64_t ao = (uint64_t)((ks * 2 * A_LBO) >> 4), bo = (uint64_t)((ks * 2 * B_LBO) >> 4);
        if (ks == 0) mma_tf32<false>(d_tmem, a_hi + ao, b + bo, id_wide);
        else mma_tf32<true>(d_tmem, a_hi + ao, b + bo, id_wide);
        mma_tf32<true>(d_tmem, a_lo + ao, b + bo, id_narrow);
    }
}
// product 1: A = x in the swizzled layout (two K blocks of 32), B = [CB_hi ; CB_lo] un-swizzled
__device__ __forceinline__ void mma_split_x(uint32_t d_tmem, uint64_t a_hi, uint64_t a_lo, uint64_t b, uint32_t id_wide, uint32_t id_narrow) {
#pragma unroll
    for (int ks = 0; ks < N / 8; ++ks) {
        const uint64_t ao = (uint64_t)(((ks >> 2) * XSW_KB + (ks & 3) * 32) >> 4), bo = (uint64_t)((ks * 2 * CBT_LBO) >> 4);
        if (ks == 0) mma_tf32<false>(d_tmem, a_hi + ao, b + bo, id_wide);
        else mma_tf32<true>(d_tmem, a_hi + ao, b + bo, id_wide);
        mma_tf32<true>(d_tmem, a_lo + ao, b + bo, id_narrow);
    }
}
// plain 3 x TF32 (three MMAs per K step into the same columns) for the two products whose accumulator is wide: reading the
// doubled accumulator back (tensor memory reads run at ~64 B/clk per SM) costs more than the third MMA
template <int KS, uint32_t A_LBO, uint32_t B_LBO>
__device__ __forceinline__ void mma_three(uint32_t d_tmem, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo, uint32_t id) {
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
        const uint64_t ao = (uint64_t)((ks * 2 * A_LBO) >> 4), bo = (uint64_t)((ks * 2 * B_LBO) >> 4);
        if (ks == 0) mma_tf32<false>(d_tmem, a_lo + ao, b_hi + bo, id);      // small terms first
        else mma_tf32<true>(d_tmem, a_lo + ao, b_hi + bo, id);
        mma_tf32<true>(d_tmem, a_hi + ao, b_lo + bo, id);
        mma_tf32<true>(d_tmem, a_hi + ao, b_hi + bo, id);
    }
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld1(uint32_t taddr, uint32_t& r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}
// 16 lanes x 2 column groups: threads 0..15 read lanes 0..15 at the given columns, threads 16..31 the same lanes 8 columns on
__device__ __forceinline__ void tmem_ld16x2_8(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.16x32bx2.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8], 8;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
#ifdef EE_TC_BOUNDED_WAIT
// development aid: trap instead of hanging when a commit never arrives
__device__ __forceinline__ void tc_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t spins = 0;; ++spins) {
        uint32_t ok;
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        if (spins > (1u << 24)) __trap();
    }
}
#else
__device__ __forceinline__ void tc_wait(uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); }
#endif

template <int kInstance>      // a template so that only the translation unit that launches it compiles it
__global__ void __launch_bounds__(kThreads, 2) hfs_tc64_kernel(const HfsArgs a, const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap ymap) {
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    const uint32_t sbase = smem_u32(tc_smem);
    const uint32_t bar = sbase, xbar = sbase + 8;             // commits of the products; arrival of the next pair's x
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tc_smem + 32);
    float* Wt = reinterpret_cast<float*>(tc_smem + OFF_W);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int quad = warp & 3;            // tensor-memory lanes 32 quad .. 32 quad + 31 (fixed by the warp's rank in its warpgroup)
    const int half = warp >> 2;           // the two warps of a quadrant share every row: each takes half of its columns

    // ---- prologue: tensor memory, barrier, constant operand images ---------------------------------------------------------
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    auto copy_pair = [&](int pair) {                         // one thread
        mbar_arrive_expect_tx(xbar, 2 * XSW_KB);
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                             sbase + OFF_XL + kb * XSW_KB),
                         "l"(reinterpret_cast<uint64_t>(&xmap)), "r"(kb * 32), "r"(pair * 2 * N), "r"(xbar)
                         : "memory");
    };
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_init(xbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
        if ((int)blockIdx.x < ((a.planes + 1) >> 1)) copy_pair((int)blockIdx.x);      // the first pair's x lands behind the prologue
    }
    auto st_pair = [&](uint32_t off_hi, uint32_t off_lo, uint32_t o, float v) {
        float hi, lo;
        split_tf32(v, hi, lo);
        *reinterpret_cast<float*>(tc_smem + off_hi + o) = hi;
        *reinterpret_cast<float*>(tc_smem + off_lo + o) = lo;
    };
    for (int e = tid; e < N * NJt; e += kThreads) {
        const int w = e / NJt, j = e % NJt;
        const float v = __ldg(a.cb + w * NJp + j);
        st_pair(OFF_CBT, OFF_CBT + CBT_BYTES, op_off(j, w, CBT_LBO, CBT_SBO), v);
        st_pair(OFF_CB, OFF_CB + CB_BYTES, op_off(w, j, CB_LBO, CB_SBO), v);
    }
    for (int e = tid; e < N * NIt; e += kThreads) {
        const int h = e / NIt, i = e % NIt;
        const float v = (i < NIp) ? __ldg(a.rb + h * NIp + i) : 0.0f;
        st_pair(OFF_RBT, OFF_RBT + RBT_BYTES, op_off(i, h, RBT_LBO, RBT_SBO), v);
        st_pair(OFF_RB, OFF_RB + RB_BYTES, op_off(h, i, RB_LBO, RB_SBO), v);
    }
    for (int e = tid; e < NIt * NJt; e += kThreads) {
        const int i = e / NJt, j = e % NJt;
        Wt[e] = (i < NIp) ? __ldg(a.w + i * NJp + j) : 0.0f;
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = tmem + ((uint32_t)(quad * 32) << 16);

    unsigned char* pA = tc_smem + OFF_XH;        // work area A
    unsigned char* pXl = tc_smem + OFF_XL;       // work area B
    // loop-invariant descriptors of the four products (used by thread 0 only)
    const uint32_t sA = sbase + OFF_XH;
    const uint64_t dXh = make_desc_sw128(sA), dXl = make_desc_sw128(sbase + OFF_XL);
    const uint64_t dCBt = make_desc(sbase + OFF_CBT, CBT_LBO, CBT_SBO);
    const uint64_t dTh = make_desc(sA, TT_LBO, TT_SBO), dTl = make_desc(sA + A_TT_LO, TT_LBO, TT_SBO);
    const uint64_t dRBt = make_desc(sbase + OFF_RBT, RBT_LBO, RBT_SBO);
    const uint64_t dRBh = make_desc(sbase + OFF_RB, RB_LBO, RB_SBO), dRBl = make_desc(sbase + OFF_RB + RB_BYTES, RB_LBO, RB_SBO);
    const uint64_t dGh = make_desc(sA + A_G_HI, G_LBO, G_SBO), dGl = make_desc(sA + A_G_LO, G_LBO, G_SBO);
    const uint64_t dVh = make_desc(sA, V_LBO, V_SBO), dVl = make_desc(sA + A_V_LO, V_LBO, V_SBO);
    const uint64_t dCB = make_desc(sbase + OFF_CB, CB_LBO, CB_SBO);
    const int npairs = (a.planes + 1) >> 1;

    // x of one pair (128 rows x 64 floats, contiguous in global memory) is brought in by the TMA engine: two tensor copies of
    // 32 floats x 128 rows with the 128 B swizzle, i.e. straight into the K-major SWIZZLE_128B operand layout, into work area B
    // (the lo operand), completion on `xbar`.  No load / store unit traffic, no registers; rows beyond the last plane are
    // zero-filled by the copy.  The split then runs in place: the thread that reads a 16 B chunk writes hi to the same offset
    // of work area A and lo back.  Chunk (row, c): offset (c / 8) * 16 KB + row * 128 + ((c % 8) ^ (row % 8)) * 16; a quarter
    // warp takes one logical chunk of 8 consecutive rows = 8 distinct physical chunks (conflict free).
    // Warp (quad, half) splits K block `half` of rows 32 quad .. 32 quad + 31 -- exactly the 4 KB of work area A that it
    // stages its part of y in at the end of the iteration, so the only hazard between one pair's output and the next pair's
    // split is inside the warp (no CTA barrier between iterations).
    auto x_off = [&](int q) {
        const int row = quad * 32 + (q >> 1) * 8 + (lane & 7), c = (q & 1) * 4 + (lane >> 3);
        return (uint32_t)half * XSW_KB + (uint32_t)row * 128u + (uint32_t)((c ^ (row & 7)) * 16);
    };

    int pair = blockIdx.x;
    uint32_t ph = 0;
#ifdef EE_TC_PROFILE
    long long prof[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tlast = 0;
    int iters = 0;
#define TC_MARK(k) do { const long long now_ = clock64(); prof[k] += now_ - tlast; tlast = now_; } while (0)
#else
#define TC_MARK(k) do { } while (0)
#endif
    uint32_t xph = 0;
#ifdef EE_TC_PROFILE
    const long long k0 = clock64();
    unsigned long long g0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
#endif
    for (; pair < npairs; pair += gridDim.x) {
#ifdef EE_TC_PROFILE
        tlast = clock64(); ++iters;
#endif
        // ---- 0. x (landed in work area B) -> hi to work area A, lo in place ----------------------------------------------------
        mbar_wait(xbar, xph); xph ^= 1u;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const uint32_t o = x_off(q);
            const float4 v = *reinterpret_cast<const float4*>(pXl + o);
            float4 hi, lo;
            split_tf32(v.x, hi.x, lo.x); split_tf32(v.y, hi.y, lo.y);
            split_tf32(v.z, hi.z, lo.z); split_tf32(v.w, hi.w, lo.w);
            *reinterpret_cast<float4*>(pA + o) = hi;
            *reinterpret_cast<float4*>(pXl + o) = lo;
        }
        fence_async_smem();
        __syncthreads();
        TC_MARK(0);
        // ---- 1. T = X CB ---------------------------------------------------------------------------------------------------
        if (tid == 0) {
            tc_fence_after();
            mma_split_x(tmem + COL_T, dXh, dXl, dCBt, idesc(128, 2 * NJt), idesc(128, NJt));
            mma_commit(bar);
        }
        const int next = pair + (int)gridDim.x;
        tc_wait(bar, ph); ph ^= 1u;
        tc_fence_after();
        const bool more = next < npairs;
        if (tid == 0 && more) copy_pair(next);               // product 1 is done with work area B: the next x lands behind the
                                                             // rest of this iteration
        TC_MARK(1);
        // ---- 2. T -> T~ (transposed: rows (plane, j), K = h); this warp's half of the 16 columns ------------------------------
        {
            uint32_t t[8], t2[8];
            tmem_ld8(lane_base + COL_T + 8 * half, t);
            tmem_ld8(lane_base + COL_T + NJt + 8 * half, t2);
            tmem_wait_ld();
            const int p = quad >> 1, h = (quad & 1) * 32 + lane;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                float hi, lo;
                split_tf32(__uint_as_float(t[jj]) + __uint_as_float(t2[jj]), hi, lo);
                const uint32_t o = op_off(p * NJt + 8 * half + jj, h, TT_LBO, TT_SBO);
                *reinterpret_cast<float*>(pA + o) = hi;
                *reinterpret_cast<float*>(pA + A_TT_LO + o) = lo;
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        TC_MARK(2);
        // ---- 3. D^T = T^T RB -------------------------------------------------------------------------------------------------
        if (tid == 0) {
            tc_fence_after();
            mma_split<N / 8, TT_LBO, RBT_LBO>(tmem + COL_D, dTh, dTl, dRBt, idesc(64, 2 * NIt), idesc(64, NIt));
            mma_commit(bar);
        }
        tc_wait(bar, ph); ph ^= 1u;
        tc_fence_after();
        TC_MARK(3);
        // ---- 4. G = W o D + cross terms; rows (plane, j), K = i; this warp's 12 of the 24 rows i ------------------------------
        if (quad < 2) {                                       // M = 64 accumulator: rows 16 q .. 16 q + 15 sit in lanes 0..15 of quadrant q
            uint32_t du[12], du2[12], xr_[2], xr2_[2];
            tmem_ld8(lane_base + COL_D + 12 * half, du);
            tmem_ld4(lane_base + COL_D + 12 * half + 8, du + 8);
            tmem_ld8(lane_base + COL_D + NIt + 12 * half, du2);
            tmem_ld4(lane_base + COL_D + NIt + 12 * half + 8, du2 + 8);
            tmem_ld1(lane_base + COL_D + R, xr_[0]);          // rows r and 2 r of D for the cross terms
            tmem_ld1(lane_base + COL_D + 2 * R, xr_[1]);
            tmem_ld1(lane_base + COL_D + NIt + R, xr2_[0]);
            tmem_ld1(lane_base + COL_D + NIt + 2 * R, xr2_[1]);
            tmem_wait_ld();
            const int j = lane & 15;
            const bool jcos = j < R;
            const int k = jcos ? j : j - (R - 1);
            const int partner = (j >= 1 && j < NJ) ? (jcos ? R - 1 + k : k) : j;
            const float dr = __shfl_sync(0xffffffffu, __uint_as_float(xr_[0]) + __uint_as_float(xr2_[0]), partner);
            const float d2r = __shfl_sync(0xffffffffu, __uint_as_float(xr_[1]) + __uint_as_float(xr2_[1]), partner);
            if (lane < 16) {
                float g[12];
#pragma unroll
                for (int ii = 0; ii < 12; ++ii) {
                    const int i = 12 * half + ii;
                    float v = Wt[i * NJt + j] * (__uint_as_float(du[ii]) + __uint_as_float(du2[ii]));
                    if (j >= 1 && j < NJ) {
                        if (i == 2 * R) v = jcos ? fmaf(-a.gamma, dr, v) : fmaf(a.gamma, dr, v);
                        if (i == R) v = jcos ? fmaf(a.gamma, d2r, v) : fmaf(-a.gamma, d2r, v);
                    }
                    g[ii] = (i < NI && j < NJ) ? v : 0.0f;
                }
                unsigned char* pGh = pA + A_G_HI;
                unsigned char* pGl = pA + A_G_LO;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float4 hi, lo;
                    split_tf32(g[4 * c + 0], hi.x, lo.x); split_tf32(g[4 * c + 1], hi.y, lo.y);
                    split_tf32(g[4 * c + 2], hi.z, lo.z); split_tf32(g[4 * c + 3], hi.w, lo.w);
                    const uint32_t o = op_off(quad * NJt + j, 12 * half + 4 * c, G_LBO, G_SBO);
                    *reinterpret_cast<float4*>(pGh + o) = hi;
                    *reinterpret_cast<float4*>(pGl + o) = lo;
                }
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        TC_MARK(4);
        // ---- 5. V = RB G -----------------------------------------------------------------------------------------------------
        if (tid == 0) {
            tc_fence_after();
            mma_three<NIt / 8, RB_LBO, G_LBO>(tmem + COL_V, dRBh, dRBl, dGh, dGl, idesc(64, 2 * NJt));
            mma_commit(bar);
        }
        tc_wait(bar, ph); ph ^= 1u;
        tc_fence_after();
        TC_MARK(5);
        // ---- 6. V -> V~: rows (plane, h), K = j; this warp's plane.  The M = 64 accumulator keeps rows 16 q .. 16 q + 15 in lanes
        //         0 .. 15 of quadrant q: the 16-lane load gives lanes 16 .. 31 the second 8 columns of the same rows ------------
        {
            uint32_t v[8];
            tmem_ld16x2_8(lane_base + COL_V + NJt * half, v);
            tmem_wait_ld();
            const int h = quad * 16 + (lane & 15), c0 = (lane >> 4) * 2;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                float4 hi, lo;
                split_tf32(__uint_as_float(v[4 * c + 0]), hi.x, lo.x); split_tf32(__uint_as_float(v[4 * c + 1]), hi.y, lo.y);
                split_tf32(__uint_as_float(v[4 * c + 2]), hi.z, lo.z); split_tf32(__uint_as_float(v[4 * c + 3]), hi.w, lo.w);
                const uint32_t o = op_off(half * N + h, 4 * (c0 + c), V_LBO, V_SBO);
                *reinterpret_cast<float4*>(pA + o) = hi;
                *reinterpret_cast<float4*>(pA + A_V_LO + o) = lo;
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        TC_MARK(6);
        // ---- 7. y = V CB^T ---------------------------------------------------------------------------------------------------
        if (tid == 0) {
            tc_fence_after();
            mma_three<NJt / 8, V_LBO, CB_LBO>(tmem + COL_Y, dVh, dVl, dCB, dCB + (uint64_t)(CB_BYTES >> 4), idesc(128, N));
            mma_commit(bar);
        }
        tc_wait(bar, ph); ph ^= 1u;
        tc_fence_after();
        TC_MARK(7);
        // ---- 8. this warp's half (32 columns) of the output rows of its quadrant.  Without `add` (or with `add` == y): tensor memory -> the
        //         128 B-swizzled layout in work area A (row per lane: 8 consecutive rows hit 8 distinct chunks) -> ONE TMA
        //         tensor store of 32 x 32 floats per warp (no load / store unit traffic to global memory).  With `add`:
        //         272 B-strided rows -> read back 4 rows x 128 B per instruction -> + add -> coalesced 128-bit stores.
        if (a.add == nullptr || a.add == a.y) {              // `add` aliasing y (the in-place accumulation of the front end's
                                                             // backward): the same store as a TMA reduction, y += staged
            const int row = quad * 32 + lane;
            unsigned char* prow = pA + half * XSW_KB + (size_t)row * 128;
#pragma unroll
            for (int blk = 0; blk < 2; ++blk) {
                uint32_t yv[16];
                tmem_ld16(lane_base + COL_Y + 32 * half + blk * 16, yv);
                tmem_wait_ld();
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    *reinterpret_cast<float4*>(prow + (((blk * 4 + c) ^ (row & 7)) * 16)) =
                        make_float4(__uint_as_float(yv[4 * c]), __uint_as_float(yv[4 * c + 1]), __uint_as_float(yv[4 * c + 2]), __uint_as_float(yv[4 * c + 3]));
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                if (a.add == nullptr)
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(reinterpret_cast<uint64_t>(&ymap)),
                                 "r"(32 * half), "r"(pair * 2 * N + quad * 32), "r"(sA + half * XSW_KB + quad * 32 * 128)
                                 : "memory");
                else
                    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                                     reinterpret_cast<uint64_t>(&ymap)),
                                 "r"(32 * half), "r"(pair * 2 * N + quad * 32), "r"(sA + half * XSW_KB + quad * 32 * 128)
                                 : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");        // this warp's 4 KB of work area A are free again
            }
            __syncwarp();
        } else
        {
            unsigned char* pw = pA + (size_t)(quad * 32) * Y_ROW + 128 * half;      // this warp's 32 rows
            unsigned char* prow = pw + (size_t)lane * Y_ROW;
#pragma unroll
            for (int blk = 0; blk < 2; ++blk) {
                uint32_t yv[16];
                tmem_ld16(lane_base + COL_Y + 32 * half + blk * 16, yv);
                tmem_wait_ld();
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    *reinterpret_cast<float4*>(prow + (blk * 16 + 4 * c) * 4) =
                        make_float4(__uint_as_float(yv[4 * c]), __uint_as_float(yv[4 * c + 1]), __uint_as_float(yv[4 * c + 2]), __uint_as_float(yv[4 * c + 3]));
            }
            __syncwarp();
            const int plane = pair * 2 + (quad >> 1);
            if (plane < a.planes) {
                const size_t base = (size_t)plane * N * N + (size_t)((quad & 1) * 32) * N + 32 * half + (lane & 7) * 4;
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int row = it * 4 + (lane >> 3);
                    float4 o = *reinterpret_cast<const float4*>(pw + (size_t)row * Y_ROW + (lane & 7) * 16);
                    if (a.add) o = f4add(o, __ldcs(reinterpret_cast<const float4*>(a.add + base + (size_t)row * N)));
                    *reinterpret_cast<float4*>(a.y + base + (size_t)row * N) = o;
                }
            }
        }
        TC_MARK(8);
        tc_fence_before();        // (every warp has drained y before the barrier that precedes the next pair's first product)
        if (a.add != nullptr && a.add != a.y) __syncthreads();      // the strided staging rows are shared between warps
    }

#ifdef EE_TC_PROFILE
    unsigned long long g1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    if (tid == 0 && blockIdx.x == 0 && iters > 0)
        printf("tc clock: %lld cycles in %llu ns\n", clock64() - k0, g1 - g0);
    if (tid == 0 && blockIdx.x == 0 && iters > 0)
        printf("tc profile (cycles per pair, %d pairs): split %lld | mma1 %lld | drainT %lld | mma2 %lld | G %lld | mma3 %lld | drainV %lld | mma4 %lld | drainY %lld\n",
               iters, prof[0] / iters, prof[1] / iters, prof[2] / iters, prof[3] / iters, prof[4] / iters, prof[5] / iters, prof[6] / iters,
               prof[7] / iters, prof[8] / iters);
#endif
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

}  // namespace hfs_tc
}  // namespace ee
