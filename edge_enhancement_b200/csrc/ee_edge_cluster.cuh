// ee_edge_cluster.cuh -- CannyFilter_step125_1 (+ blend) backward for images that do not fit one CTA
// (ImageNet 224 px): ONE THREAD-BLOCK CLUSTER PER IMAGE, halo rows exchanged through distributed shared memory.
// EXPERIMENTAL / opt-in (see the measurement note below).
//
// The strip kernels of ee_edge_fast.cuh recompute a halo (4 rows above and below, one float4 group left and right
// of every tile: ~1.4x the arithmetic and the L2 reads at 224 px).  Here the CS CTAs of a cluster own TH = H / CS
// consecutive full-width rows each and compute every stage ONLY on their own rows; the one row above / below that
// a 3x3 stage needs at a CTA boundary is read straight out of the neighbouring CTA's shared memory
// (cluster.map_shared_rank -> ld.shared::cluster).  Stages are separated by cluster-wide barriers, so the region
// aliasing of the single-CTA kernel (S -> A, Bl -> GB) stays race-free.  Every thread owns exactly one chunk of R rows
// x 4 columns per stage (blockDim = W/4 x TH/R), as in the whole-image ("EVEN") kernels: no row guards, no loops.
//
// Same canonical arithmetic, same bits as the other implementations (tests/test_gpu_parity.py compares them).
#pragma once
#include <cooperative_groups.h>

#include "ee_edge_fast.cuh"

namespace ee {

namespace cg = cooperative_groups;

// Where the rows of one plane live for this CTA: own rows [0, TH) locally, the row above row 0 and the row below row
// TH-1 behind `up` / `dn` (a neighbour's plane through DSMEM; at an image border the own border row for
// replicate-extended planes, or nullptr for zero-extended ones).  Pointers address column 0 of a row (pads at -1 / W).
struct ClRows {
    const float* own;
    const float* up;
    const float* dn;
};

// Measured on B200 (512x3x224x224, profiles/README.md): 4.02 TB/s with cluster.sync() between the stages -- the same as
// the strip kernels (4.05 TB/s), whose redundant halo work costs about what the five cluster barriers cost here
// (barrier + membar = 37 % of the stall samples).  A split barrier (barrier.cluster.arrive.release right after a
// stage's stores, .wait.acquire only in the threads that touch a neighbour's row, interior chunks first) was SLOWER
// (3.26 TB/s: the unaligned per-thread waits serialise).  The strip kernels therefore stay the default; this kernel is
// selected with ee_set_tuning(.., .., staging = 5) and kept bit-identical by tests/test_gpu_parity.py.
template <int R, int TH>
__device__ __forceinline__ const float* cl_row(const ClRows& p, int i, int lr, int Wp) {
    // i is the unrolled slot index: only slots 0 and R+1 can fall outside the own rows
    if (i == 0 && lr < 0) return p.up;
    if (i == R + 1 && lr >= TH) return p.dn;
    return p.own + lr * Wp;
}

template <int NC, bool BLEND, int R, int W, int TH, int CS>
__global__ void __launch_bounds__((W / 4) * (TH / R) <= 256 ? 256 : (((W / 4) * (TH / R) + 31) / 32) * 32, 2)
edge_bwd_step125_cluster(const FastArgs a) {
    extern __shared__ __align__(16) float smem[];
    constexpr int DIVM = (NC == 1) ? 0 : (NC == 3 ? 1 : 2);
    constexpr int Wp = W + kPadW, GX = W / 4, RY = TH / R, H = TH * CS;
    static_assert(TH % R == 0 && W % 4 == 0, "cluster tiles are whole chunks");
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int b = blockIdx.x / CS;
    const int C = NC ? NC : a.e.C;
    const int tx = threadIdx.x % GX, ty = threadIdx.x / GX;
    const bool active = ty < RY;
    const size_t hw = (size_t)H * W;
    const int r0 = rank * TH;                                  // first image row of this CTA
    const int lc = tx * 4, ra = ty * R;                        // this thread's chunk: local rows [ra, ra+R), columns [lc, lc+4)
    const bool left = (tx == 0), right = (tx == GX - 1);
    const bool first = (rank == 0), last = (rank == CS - 1);
    const bool want_gx = (a.e.g_x != nullptr);

    float* R1 = smem;                      // S  -> A
    float* R2 = R1 + TH * Wp;              // Bl -> GB
    float* R3 = R2 + TH * Wp;              // Bv
    // the same regions in the neighbouring CTAs (generic pointers into distributed shared memory)
    const float* R1u = first ? nullptr : cluster.map_shared_rank(R1, rank - 1) + (TH - 1) * Wp;
    const float* R2u = first ? nullptr : cluster.map_shared_rank(R2, rank - 1) + (TH - 1) * Wp;
    const float* R3u = first ? nullptr : cluster.map_shared_rank(R3, rank - 1) + (TH - 1) * Wp;
    const float* R1d = last ? nullptr : cluster.map_shared_rank(R1, rank + 1);
    const float* R2d = last ? nullptr : cluster.map_shared_rank(R2, rank + 1);
    const float* R3d = last ? nullptr : cluster.map_shared_rank(R3, rank + 1);

    // ---- stage 0: S = sum over channels of the own rows (replicate pad columns) -------------------------------
    if (active) {
        const float* px = a.e.x + (size_t)b * C * hw + (size_t)(r0 + ra) * W + lc;
        float4 acc[R];
        if (NC == 3) {
            float4 v1[R], v2[R];
#pragma unroll
            for (int i = 0; i < R; ++i) {
                acc[i] = __ldg(reinterpret_cast<const float4*>(px + i * W));
                v1[i] = __ldg(reinterpret_cast<const float4*>(px + i * W + hw));
                v2[i] = __ldg(reinterpret_cast<const float4*>(px + i * W + 2 * hw));
            }
#pragma unroll
            for (int i = 0; i < R; ++i) acc[i] = f4add(f4add(acc[i], v1[i]), v2[i]);
        } else {
#pragma unroll
            for (int i = 0; i < R; ++i) acc[i] = __ldg(reinterpret_cast<const float4*>(px + i * W));
            for (int c = 1; c < C; ++c) {
#pragma unroll
                for (int i = 0; i < R; ++i) acc[i] = f4add(acc[i], __ldg(reinterpret_cast<const float4*>(px + i * W + (size_t)c * hw)));
            }
        }
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const float o[4] = {acc[i].x, acc[i].y, acc[i].z, acc[i].w};
            st_plane(R1 + (ra + i) * Wp + kPadL + lc, o, left, right, o[0], o[3]);
        }
    }
    // L2 bulk prefetch of the A/Bv stage's operands (own rows of base and g_out), after the x loads are issued
    if (C <= 32) {
        if (BLEND) {
            if (threadIdx.x < 32) prefetch_rows(a.e.base, b, C, H, W, r0, r0 + TH, threadIdx.x);
            else if (threadIdx.x < 64) prefetch_rows(a.e.g_in, b, C, H, W, r0, r0 + TH, threadIdx.x - 32);
        } else if (threadIdx.x == 0) {
            prefetch_rows(a.e.g_in, b, 1, H, W, r0, r0 + TH, 0);
        }
    }
    cluster.sync();

    // ---- stage 1: Bl = blur(S), replicate extension (at an image border the row above row 0 is row 0 itself) ------
    if (active) {
        const ClRows S = {R1, first ? R1 : R1u, last ? R1 + (TH - 1) * Wp : R1d};
        const float c0 = a.e.c0, c1 = a.e.c1, c2 = a.e.c2;
        float P[3][4], Q[3][4];
#pragma unroll
        for (int i = 0; i < R + 2; ++i) {
            gauss_partials(ld_win(cl_row<R, TH>(S, i, ra - 1 + i, Wp) + kPadL + lc), c0, c1, c2, P[i % 3], Q[i % 3]);
            if (i >= 2) {
                float o[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) o[k] = (P[(i - 2) % 3][k] + Q[(i - 1) % 3][k]) + P[i % 3][k];
                st_plane(R2 + (ra + i - 2) * Wp + kPadL + lc, o, left, right, o[0], o[3]);
            }
        }
    }
    cluster.sync();

    // ---- stage 2: A / Bv = dL/dSgx, dL/dSgy on the own rows (zero pad columns), g_base ----------------------------
    if (active) {
        const ClRows Bl = {R2, first ? R2 : R2u, last ? R2 + (TH - 1) * Wp : R2d};
        const float* base_b = a.e.base + (size_t)b * C * hw;
        const float* gin_b = a.e.g_in + (size_t)b * (BLEND ? C : 1) * hw;
        float* gbase_b = a.e.g_base ? a.e.g_base + (size_t)b * C * hw : nullptr;
        float D[3][4], V[3][4];
#pragma unroll
        for (int i = 0; i < R + 2; ++i) {
            sobel_partials(ld_win(cl_row<R, TH>(Bl, i, ra - 1 + i, Wp) + kPadL + lc), D[i % 3], V[i % 3]);
            if (i >= 2) {
                const int lr = ra + i - 2, q = lr * Wp + kPadL + lc;
                bwd_abv_row<NC, BLEND, false, DIVM>(a, D[(i - 2) % 3], D[(i - 1) % 3], D[i % 3], V[(i - 2) % 3], V[i % 3],
                                                    (r0 + lr) * W + lc, true, left, right, base_b, gin_b, gbase_b, hw, want_gx,
                                                    R1 + q, R3 + q);
            }
        }
    }
    cluster.sync();
    if (!want_gx) return;        // uniform over the cluster; nobody reads this CTA's planes any more

    // ---- stage 3: GB = fold(Sobel^T(A, Bv)) on the own rows; ring rows only at the image borders ------------------
    const AdjBorder bd = {left, right};
    const bool ring = left || right;
    const int gra = r0 + ra;                                   // image row of the chunk (adj_chunk_even works in image rows)
    if (active) {
        const ClRows A = {R1, R1u, R1d}, Bv = {R3, R3u, R3d};
        float HA[3][4], HB[3][4], HAr[3], HBr[3];
        auto loadp = [&](int i, int rin, bool valid) {
            if (valid) {
                const int lr = rin - r0;
                const float* pa = ((i % (R + 2)) >= 1 && (i % (R + 2)) <= R) ? A.own + lr * Wp : (lr < 0 ? A.up : (lr >= TH ? A.dn : A.own + lr * Wp));
                const float* pb = ((i % (R + 2)) >= 1 && (i % (R + 2)) <= R) ? Bv.own + lr * Wp : (lr < 0 ? Bv.up : (lr >= TH ? Bv.dn : Bv.own + lr * Wp));
                sobel_adj_partials(ld_win(pa + kPadL + lc), ld_win(pb + kPadL + lc), bd, HA[i % 3], HB[i % 3], HAr[i % 3], HBr[i % 3]);
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
            st_plane(R2 + (row - r0) * Wp + kPadL + lc, o, left, right, 0.0f, 0.0f);
        };
        adj_chunk_even<R>(gra, H, loadp, combine, store);
    }
    cluster.sync();

    // ---- stage 4: g_s = fold(Gauss^T(GB)) on the own rows -> every channel of g_x -------------------------------
    if (active) {
        const ClRows GB = {R2, R2u, R2d};
        const float c0 = a.e.c0, c1 = a.e.c1, c2 = a.e.c2;
        float* gx_b = a.e.g_x + (size_t)b * C * hw;
        float P[3][4], Q[3][4], Pr[3], Qr[3];
        auto loadp = [&](int i, int rin, bool valid) {
            if (valid) {
                const int lr = rin - r0;
                const float* pg = ((i % (R + 2)) >= 1 && (i % (R + 2)) <= R) ? GB.own + lr * Wp : (lr < 0 ? GB.up : (lr >= TH ? GB.dn : GB.own + lr * Wp));
                gauss_adj_partials(ld_win(pg + kPadL + lc), bd, c0, c1, c2, P[i % 3], Q[i % 3], Pr[i % 3], Qr[i % 3]);
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
            float* pg = gx_b + (size_t)row * W + lc;
            if (NC) {
#pragma unroll
                for (int c = 0; c < (NC ? NC : 1); ++c) __stcs(reinterpret_cast<float4*>(pg + c * hw), v);
            } else {
                for (int c = 0; c < C; ++c) __stcs(reinterpret_cast<float4*>(pg + c * hw), v);
            }
        };
        adj_chunk_even<R>(gra, H, loadp, combine, store);
    }
    cluster.sync();              // keep this CTA's planes alive until the neighbours have read their halo rows
}

}  // namespace ee
