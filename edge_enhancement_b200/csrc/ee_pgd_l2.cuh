// ee_pgd_l2.cuh -- TRADES PGD-L2 step (utils/attacks.py:391-399 with l2_norm / squared_l2_norm of :360-366) as ONE pass
// over HBM: 16 B/element (read g, x, x0; write x').
//
//     g  /= sqrt(mean(g^2)) + 1e-8          per sample (mean, not sum)
//     xa  = x + step * g ;  d = xa - x0
//     if sqrt(mean(d^2)) > eps:  d *= eps / sqrt(mean(d^2))
//     x'  = clamp(x0 + d, 0, 1)
//
// The two per-sample norms are sequentially dependent, so the sample has to stay on chip between them.  One sample is
// owned by a thread-block CLUSTER of K CTAs (K = 1 up to 3 x 64 x 64, 4 at 128 px, 8 from 224 px); CTA k stages its slice of
// g in shared memory with TMA bulk copies (cp.async.bulk + mbarrier, issued by one thread at kernel start together with an
// L2 prefetch of the x slice), overwrites it with d in the second phase, and the cluster exchanges the K partial sums
// through distributed shared memory.  With K = 1 the x0 slice is staged the same way (96 KB per CTA at 3 x 64 x 64, two
// CTAs per SM); with K > 1 only g / d stay on chip (<= 75 KB per CTA at 224 px, three CTAs per SM) and the third phase
// re-reads x0 through L2, where the second phase left it microseconds earlier.
//
// Canonical reduction order (shared with oracle/ee_oracle.c, rms_cluster): inside a slice thread t of T accumulates the
// float4 words t, t + T, ... element by element with fmaf; warp-shuffle tree (strides 16..1); the T/32 warp sums padded
// to 32 with zeros and the same tree; then the K slice sums left to right.  T = 512, or 128 for slices under 1024 words.
#pragma once
#include <cooperative_groups.h>

#include "ee_attack.cuh"
#include "ee_edge_stream.cuh"      // mbarrier / bulk-copy helpers

namespace ee {

constexpr int kL2Threads = 512;            // launch bound; small slices run 128 threads
constexpr int kL2OneCta4 = 3072;           // up to 12288 floats (48 KB per plane) one CTA keeps g AND x0: 96 KB, 2 CTAs per SM
constexpr int kL2TargetSlice4 = 4096;      // clusters: 64 KB of g per CTA when K <= 8 allows it
constexpr int kL2MaxCluster = 8;           // portable cluster size
constexpr int kL2Header = 128;             // mbarrier, warp partials, cluster slots

struct L2Plan { int K; int slice4; int threads; int stage_x0; size_t smem; };

// false: the sample does not fit (or n_per % 4 != 0) -> the three-pass kernel of ee_attack.cuh.
// oracle/ee_oracle.c (l2_cluster_plan) restates this function: keep the two in step.
inline bool pgd_l2_plan(int64_t n_per, L2Plan& pl) {
    if (n_per <= 0 || (n_per & 3)) return false;
    const int64_t n4 = n_per >> 2;
    int K = 1;
    if (n4 > kL2OneCta4)
        while (K < kL2MaxCluster && (n4 + K - 1) / K > kL2TargetSlice4) K <<= 1;
    if (n4 > kL2OneCta4 && K == 1) K = 2;
    const int64_t s4 = (n4 + K - 1) / K;
    pl.stage_x0 = (K == 1);
    const size_t smem = kL2Header + (size_t)s4 * 16 * (pl.stage_x0 ? 2 : 1);
    if (smem > (size_t)227 * 1024) return false;
    pl.K = K; pl.slice4 = (int)s4; pl.smem = smem;
    pl.threads = (s4 >= 1024) ? kL2Threads : 128;
    return true;
}

__device__ __forceinline__ float l2_block_sum(float v, float* sh) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) v = v + __shfl_down_sync(0xffffffffu, v, s);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();                 // protect sh from the previous use
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    float t = (lane < (int)(blockDim.x >> 5)) ? sh[lane] : 0.0f;
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) t = t + __shfl_down_sync(0xffffffffu, t, s);
    return __shfl_sync(0xffffffffu, t, 0);
}

static __global__ void __launch_bounds__(kL2Threads) pgd_l2_cluster_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                                          const float* __restrict__ x0, float* __restrict__ out,
                                                                          int64_t n_per, int slice4, int K, int stage_x0, float step, float eps) {
    namespace cg = cooperative_groups;
    extern __shared__ __align__(128) unsigned char l2_smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(l2_smem);
    float* red = reinterpret_cast<float*>(l2_smem + 16);          // 16 warp partials
    float* slots = reinterpret_cast<float*>(l2_smem + 96);        // [0]: sum g^2, [1]: sum d^2 of this CTA's slice
    float4* pg = reinterpret_cast<float4*>(l2_smem + kL2Header);  // g slice, then d slice
    float4* px0 = pg + slice4;
    cg::cluster_group cl = cg::this_cluster();
    const int rank = (K > 1) ? (int)cl.block_rank() : 0;
    const int64_t sample = blockIdx.x / K;
    const int64_t n4 = n_per >> 2;
    const int64_t lo4 = (int64_t)rank * slice4;
    const int cnt = (int)max((int64_t)0, min(lo4 + slice4, n4) - lo4);
    const float4* gb = reinterpret_cast<const float4*>(g + sample * n_per) + lo4;
    const float4* xb = reinterpret_cast<const float4*>(x + sample * n_per) + lo4;
    const float4* x0b = reinterpret_cast<const float4*>(x0 + sample * n_per) + lo4;
    float4* ob = reinterpret_cast<float4*>(out + sample * n_per) + lo4;

    if (threadIdx.x == 0) {
        mbar_init(smem_u32(bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (cnt > 0) {
            const uint32_t bytes = (uint32_t)cnt * 16u;
            mbar_arrive_expect_tx(smem_u32(bar), (stage_x0 ? 2u : 1u) * bytes);
            for (uint32_t off = 0; off < bytes; off += 32768u) {
                const uint32_t c = min(32768u, bytes - off);
                bulk_g2s(smem_u32(pg) + off, reinterpret_cast<const char*>(gb) + off, c, smem_u32(bar));
                if (stage_x0) bulk_g2s(smem_u32(px0) + off, reinterpret_cast<const char*>(x0b) + off, c, smem_u32(bar));
                else l2_prefetch_bulk(reinterpret_cast<const char*>(x0b) + off, c);
                l2_prefetch_bulk(reinterpret_cast<const char*>(xb) + off, c);
            }
        }
    }
    __syncthreads();
    if (cnt > 0) mbar_wait(smem_u32(bar), 0);

    auto cluster_total = [&](float part, int which) {
        if (K == 1) return part;
        if (threadIdx.x == 0) slots[which] = part;
        cl.sync();
        float tot = *cl.map_shared_rank(slots + which, 0);
        for (int k = 1; k < K; ++k) tot = tot + *cl.map_shared_rank(slots + which, k);
        return tot;
    };

    // ---- sqrt(mean(g^2)) + 1e-8                                                                  attacks.py:391, :360-366
    float acc = 0.0f;
    const int T = (int)blockDim.x;
    for (int i = threadIdx.x; i < cnt; i += T) {
        const float4 v = pg[i];
        acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
    }
    const float gn = sqrtf(cluster_total(l2_block_sum(acc, red), 0) / (float)n_per) + 1e-8f;

    // ---- xa = x + step * g / gn ; d = xa - x0 ; sqrt(mean(d^2))                                  attacks.py:392-395
    acc = 0.0f;
    for (int i = threadIdx.x; i < cnt; i += T) {
        const float4 xv = __ldcs(xb + i), gv = pg[i], zv = stage_x0 ? px0[i] : __ldg(x0b + i);
        float4 d;
        d.x = (xv.x + step * (gv.x / gn)) - zv.x; d.y = (xv.y + step * (gv.y / gn)) - zv.y;
        d.z = (xv.z + step * (gv.z / gn)) - zv.z; d.w = (xv.w + step * (gv.w / gn)) - zv.w;
        acc = fmaf(d.x, d.x, acc); acc = fmaf(d.y, d.y, acc); acc = fmaf(d.z, d.z, acc); acc = fmaf(d.w, d.w, acc);
        pg[i] = d;                                   // same thread re-reads it below
    }
    const float dn = sqrtf(cluster_total(l2_block_sum(acc, red), 1) / (float)n_per);
    const bool cond = dn > eps;                                                                   // attacks.py:396
    const float scale = eps / dn;                                                                 // :397

    // ---- x' = clamp(x0 + d, 0, 1)                                                                attacks.py:398-399
    for (int i = threadIdx.x; i < cnt; i += T) {
        float4 d = pg[i];
        const float4 zv = stage_x0 ? px0[i] : __ldcs(x0b + i);
        if (cond) { d.x = d.x * scale; d.y = d.y * scale; d.z = d.z * scale; d.w = d.w * scale; }
        __stcs(ob + i, make_float4(minn(maxn(zv.x + d.x, 0.0f), 1.0f), minn(maxn(zv.y + d.y, 0.0f), 1.0f),
                                   minn(maxn(zv.z + d.z, 0.0f), 1.0f), minn(maxn(zv.w + d.w, 0.0f), 1.0f)));
    }
    if (K > 1) cl.sync();            // keep the slots alive until every CTA of the cluster has read them
}

}  // namespace ee
