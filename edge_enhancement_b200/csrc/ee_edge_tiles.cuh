// ee_edge_tiles.cuh -- CannyFilter_step125_1 (+ blend) backward for WIDE images (ImageNet 224 / 288 px) on
// chunk-aligned tiles: the guard-free "one chunk per thread" code of the whole-image kernels, applied to tiles.
//
// A tile is TH x TW output pixels (TH % 4 == 0, TW = 56) of one image; its three planes cover the tile plus a halo of
// 4 rows and one float4 group on every side, clipped to the image: at most 64 x 64 = 16 x 16 chunks of 4 rows x 4
// columns, ONE PER THREAD, in every stage.  Every stage is computed on the WHOLE plane (rows [r0-4, r1+4)), although
// the dependency cone only needs r0-4 / -3 / -2 / -1 / 0: the few extra rows are "garbage" rows -- finite values
// computed from clamped row indices that no valid output ever reads -- and in exchange every chunk is complete and
// 4-aligned in image coordinates (H % 4 == 0), so that
//   * there are no chunk loops and no "is this row inside my chunk" guards (as in the HT-specialised kernels),
//   * an image border (replicate rows, ring-row fold of the adjoint) can only coincide with a chunk border.
// Replaces the guarded strip path of ee_edge_fast.cuh for these shapes (same bits: tests/test_gpu_parity.py).
//
// TMA = true (C == 3): the three channel tiles of x -- the plane INCLUDING its pad columns, i.e. a box of PW + 8
// columns x TH + 8 rows starting at column cs - 4 -- are staged by the TMA engine (cp.async.bulk.tensor.3d over a
// [B*C, H, W] tensor map, out-of-bounds columns zero-filled) straight into the three plane regions, one elected
// thread issuing the copies and everybody waiting on an mbarrier; the channel sum then runs shared -> shared in
// place.  TMA = false: 128-bit LDGs summed in registers (the TMA engine cannot add channels in flight).
#pragma once
#include <cuda.h>
#include <cuda/barrier>

#include "ee_edge_fast.cuh"

namespace ee {

template <int NC, bool BLEND, int R, int PW, bool TMA = false>
__global__ void __launch_bounds__(256, 3) edge_bwd_step125_tiles(const FastArgs a, const __grid_constant__ CUtensorMap x_map) {
    extern __shared__ __align__(128) float smem_tiles[];       // TMA destinations must be 128-byte aligned
    float* smem = smem_tiles;
    constexpr int DIVM = (NC == 1) ? 0 : (NC == 3 ? 1 : 2);
    constexpr int Wp = PW + kPadW, GXT = PW / 4;               // plane row stride; thread columns (16 for PW = 64)
    static_assert(R == 4 && PW % 4 == 0 && (256 % GXT) == 0, "16 x 16 chunks of 4 x 4 pixels");
    constexpr int RYT = 256 / GXT;
    const int H = a.e.H, W = a.e.W, C = NC ? NC : a.e.C;
    const int b = blockIdx.x / a.e.tiles_per_img;
    const int tq = blockIdx.x - b * a.e.tiles_per_img;
    const int ti = tq / a.tiles_x, tj = tq - ti * a.tiles_x;
    const int r0 = ti * a.e.TH, r1 = min(r0 + a.e.TH, H);
    const int c0 = tj * a.TW, c1 = min(c0 + a.TW, W);
    const int a_lo = max(r0 - 4, 0), a_hi = min(r1 + 4, H);    // plane rows (multiples of 4)
    const int cs = max(c0 - 4, 0), ce = min(c1 + 4, W);        // plane columns
    const int n_ch = (a_hi - a_lo) >> 2, Gt = (ce - cs) >> 2;
    const int tx = threadIdx.x % GXT, ty = threadIdx.x / GXT;
    const bool active = (tx < Gt) && (ty < n_ch);
    const int lc = tx * 4, col = cs + lc;                      // plane / image column of this thread's group
    const int ra = a_lo + ty * R;                              // image row of this thread's chunk
    const bool p_left = (lc == 0), p_right = (lc + 4 == ce - cs);          // plane edges: the writer fills the pad column
    const bool want_gx = (a.e.g_x != nullptr);
    const size_t hw = (size_t)H * W;
    (void)RYT;

    float* R1 = smem;                          // S  -> A
    float* R2 = R1 + (a.e.TH + 8) * Wp;        // Bl -> GB
    float* R3 = R2 + (a.e.TH + 8) * Wp;        // Bv
    auto prow = [&](int r) { return (min(max(r, a_lo), a_hi - 1) - a_lo) * Wp + kPadL + lc; };   // clamped plane row

    // ---- stage 0: S = channel sum of the plane rows (pad columns: replicate; only meaningful at image edges) -----
    if constexpr (TMA) {
        static_assert(!TMA || NC == 3, "the TMA staging path is built for C = 3");
        namespace cde = cuda::device::experimental;
#pragma nv_diag_suppress static_var_with_dynamic_init
        __shared__ cuda::barrier<cuda::thread_scope_block> bar;
        if (threadIdx.x == 0) {
            init(&bar, 256);
            cde::fence_proxy_async_shared_cta();
        }
        __syncthreads();
        cuda::barrier<cuda::thread_scope_block>::arrival_token tok;
        if (threadIdx.x == 0) {
            // box = (PW + 8) columns x (TH + 8) rows x 1 plane; column cs - 4 lands on plane offset 0, so image column cs
            // sits at kPadL and the pad columns hold the real neighbours (zeros outside the image)
            cde::cp_async_bulk_tensor_3d_global_to_shared(R1, &x_map, cs - kPadL, a_lo, b * 3 + 0, bar);
            cde::cp_async_bulk_tensor_3d_global_to_shared(R2, &x_map, cs - kPadL, a_lo, b * 3 + 1, bar);
            cde::cp_async_bulk_tensor_3d_global_to_shared(R3, &x_map, cs - kPadL, a_lo, b * 3 + 2, bar);
            tok = cuda::device::barrier_arrive_tx(bar, 1, 3u * (uint32_t)((a.e.TH + 8) * Wp * sizeof(float)));
        } else {
            tok = bar.arrive();
        }
        bar.wait(std::move(tok));
        if (active) {
#pragma unroll
            for (int i = 0; i < R; ++i) {
                const int q = prow(ra + i);
                const float4 v0 = *reinterpret_cast<const float4*>(R1 + q), v1 = *reinterpret_cast<const float4*>(R2 + q),
                             v2 = *reinterpret_cast<const float4*>(R3 + q);
                const float4 sum = f4add(f4add(v0, v1), v2);
                const float o[4] = {sum.x, sum.y, sum.z, sum.w};
                st_plane(R1 + q, o, p_left, p_right, o[0], o[3]);
            }
        }
    } else if (active) {
        const float* px = a.e.x + (size_t)b * C * hw + (size_t)ra * W + col;
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
            st_plane(R1 + prow(ra + i), o, p_left, p_right, o[0], o[3]);
        }
    }
    // L2 bulk prefetch of the A/Bv stage's operands: one segment per (tensor, channel, needed row) of the plane columns
    if (BLEND && C <= 32) {
        const int q_lo = max(r0 - 2, 0), q_hi = min(r1 + 2, H);
        const int nrows = q_hi - q_lo, seg = (ce - cs) * (int)sizeof(float);
        for (int i = threadIdx.x; i < 2 * C * nrows; i += 256) {
            const int t = i / (C * nrows), rem = i - t * (C * nrows), c = rem / nrows, r = q_lo + rem - c * nrows;
            const float* src = (t ? a.e.g_in : a.e.base) + ((size_t)b * C + c) * hw + (size_t)r * W + cs;
            l2_prefetch_bulk(src, (uint32_t)seg);
        }
    }
    __syncthreads();

    // ---- stage 1: Bl = blur(S) on the plane rows; row indices clamped to the plane (= replicate at an image border) ----
    if (active) {
        const float c0g = a.e.c0, c1g = a.e.c1, c2g = a.e.c2;
        float P[3][4], Q[3][4];
#pragma unroll
        for (int i = 0; i < R + 2; ++i) {
            gauss_partials(ld_win(R1 + prow(ra - 1 + i)), c0g, c1g, c2g, P[i % 3], Q[i % 3]);
            if (i >= 2) {
                float o[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) o[k] = (P[(i - 2) % 3][k] + Q[(i - 1) % 3][k]) + P[i % 3][k];
                st_plane(R2 + prow(ra + i - 2), o, p_left, p_right, o[0], o[3]);
            }
        }
    }
    __syncthreads();

    // ---- stage 2: A / Bv on the plane rows; rows outside [r0-2, r1+2) are not needed: zeros, no global loads -------
    if (active) {
        const float* base_b = a.e.base + (size_t)b * C * hw;
        const float* gin_b = a.e.g_in + (size_t)b * (BLEND ? C : 1) * hw;
        float* gbase_b = a.e.g_base ? a.e.g_base + (size_t)b * C * hw : nullptr;
        const int q_lo = want_gx ? r0 - 2 : r0, q_hi = want_gx ? r1 + 2 : r1;
        const bool col_in = (col >= c0 && col < c1);
        float D[3][4], V[3][4];
#pragma unroll
        for (int i = 0; i < R + 2; ++i) {
            sobel_partials(ld_win(R2 + prow(ra - 1 + i)), D[i % 3], V[i % 3]);
            if (i >= 2) {
                const int rout = ra + i - 2;
                float* pa = R1 + prow(rout);
                float* pb = R3 + prow(rout);
                if (rout >= q_lo && rout < q_hi) {
                    bwd_abv_row<NC, BLEND, false, DIVM>(a, D[(i - 2) % 3], D[(i - 1) % 3], D[i % 3], V[(i - 2) % 3], V[i % 3],
                                                        rout * W + col, col_in && rout >= r0 && rout < r1, p_left, p_right, base_b,
                                                        gin_b, gbase_b, hw, want_gx, pa, pb);
                } else if (want_gx) {
                    const float z[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                    st_plane(pa, z, p_left, p_right, 0.0f, 0.0f);
                    st_plane(pb, z, p_left, p_right, 0.0f, 0.0f);
                }
            }
        }
    }
    if (!want_gx) return;
    __syncthreads();

    // ---- stage 3: GB = fold(Sobel^T(A, Bv)) on the plane rows (ring rows / columns only at image borders) ---------
    const AdjBorder bd = {col == 0, col + 4 == W};
    const bool ring = bd.left || bd.right;
    if (active) {
        float HA[3][4], HB[3][4], HAr[3], HBr[3];
        auto loadp = [&](int i, int rin, bool valid) {
            if (valid) {
                const int q = prow(rin);
                sobel_adj_partials(ld_win(R1 + q), ld_win(R3 + q), bd, HA[i % 3], HB[i % 3], HAr[i % 3], HBr[i % 3]);
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
        auto store = [&](int row, const float (&o)[4]) { st_plane(R2 + prow(row), o, p_left, p_right, 0.0f, 0.0f); };
        adj_chunk_even<R>(ra, H, loadp, combine, store);
    }
    __syncthreads();

    // ---- stage 4: g_s = fold(Gauss^T(GB)) on the tile's own rows and columns -> every channel of g_x -------------
    const int ro = r0 + ty * R;                                // output chunk of this thread
    if (tx < Gt && ro < r1 && col >= c0 && col < c1) {
        const float c0g = a.e.c0, c1g = a.e.c1, c2g = a.e.c2;
        float* gx_b = a.e.g_x + (size_t)b * C * hw;
        float P[3][4], Q[3][4], Pr[3], Qr[3];
        auto loadp = [&](int i, int rin, bool valid) {
            if (valid) {
                gauss_adj_partials(ld_win(R2 + prow(rin)), bd, c0g, c1g, c2g, P[i % 3], Q[i % 3], Pr[i % 3], Qr[i % 3]);
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
            float* pg = gx_b + (size_t)row * W + col;
            if (NC) {
#pragma unroll
                for (int c = 0; c < (NC ? NC : 1); ++c) __stcs(reinterpret_cast<float4*>(pg + c * hw), v);
            } else {
                for (int c = 0; c < C; ++c) __stcs(reinterpret_cast<float4*>(pg + c * hw), v);
            }
        };
        adj_chunk_even<R>(ro, H, loadp, combine, store);
    }
}

}  // namespace ee
