// ee_hfs.cuh -- HighFreqSuppress (utils/core.py:15-55), the square low-pass in front of every *_EE model, as ONE
// kernel per direction instead of torch.fft's rfft2 -> mask multiply -> irfft2 (three library passes over complex
// intermediates; measured 621 us forward at 4096x3x64x64, against 91 us for the fused edge forward next to it).
//
// The low-pass keeps the frequency rows k1 in [-r, r-1] and the one-sided columns k2 in [0, r-1]; with the C2R inverse's
// treatment of the k2 = 0 column it is the real SYMMETRIC operator (tests/test_host_logic.py)
//     y = A x Qc^T - Bm x Qs^T ,
// whose circulant factors have rank 2r+1 / 2 (rows) and 2r-1 / 2r-2 (columns).  With the real Fourier bases
//     CB[w][j] : 1, cos(k th_w) (k = 1..r-1), sin(k th_w) (k = 1..r-1)                 (NJ = 2r-1 columns)
//     RB[h][i] : 1, cos(k th_h) (k = 1..r),   sin(k th_h) (k = 1..r)                   (NI = 2r+1 columns)
// it is five small dense products per N x N plane, all in shared memory (fp32 FFMA, ~38 multiply-adds per element):
//     T = x CB      D = RB^T T      G = W o D (+ four cross terms of the -r row)      V = RB G      y = V CB^T
// The operator is self-adjoint, so the backward is the same kernel applied to the upstream gradient.
//
// Work decomposition: P planes per 256-thread CTA (256 / P threads each), whole planes resident in shared memory.
// The two large products use 4 x 4 register tiles whose 4 rows (resp. columns) are INTERLEAVED with stride N/4, so
// that the threads of a warp touch consecutive shared-memory rows (row strides = 4 mod 32 floats: every 128-bit load of
// a quarter warp hits 8 distinct bank groups).  Every sum is one fmaf chain in ascending index order: the C oracle
// (oracle/ee_oracle.c) evaluates the same chains on the same tables, so kernel and oracle agree bit for bit.
//
// The CTAs are PERSISTENT (grid = resident CTAs): a CTA loops over groups of P planes, and as soon as stage 1 has
// consumed a group's planes it starts the asynchronous copy (cp.async, LDGSTS) of the NEXT group's planes into the same
// buffer, so the global-load latency -- 43 % of the stall samples of the first, non-persistent version -- overlaps
// with stages 2-5 (two thirds of the arithmetic).
#pragma once
#include <cuda_pipeline.h>

#include <type_traits>

#include "ee_device.cuh"

namespace ee {

template <int N, int R>
struct HfsDims {
    static constexpr int NJ = 2 * R - 1, NI = 2 * R + 1;
    static constexpr int NJp = (NJ + 3) / 4 * 4, NIp = (NI + 3) / 4 * 4;
    static constexpr int XS = (N % 32 == 0) ? N + 4 : ((N + 31) / 32 * 32 + 4);   // == 4 (mod 32), >= N
    static constexpr int JS = NJp + 4;        // row stride of CB, T, V
    static constexpr int IS = NIp + 4;        // row stride of RB
    // floats of shared memory: tables (CB, RB, W) + per plane (X, T, V, D, G)
    static constexpr int kTables = N * JS + N * IS + NIp * NJp;
    static constexpr int kPlane = N * XS + N * JS + 2 * NIp * NJp;      // V reuses T's region (T is dead after stage 2)
};

struct HfsArgs {
    const float* x;      // [planes][N][N]
    float* y;            // [planes][N][N]
    const float* add;    // optional [planes][N][N]: y = H x + add (the backward accumulates into the edge-path gradient)
    const float* cb;     // [N][NJp]   column bases (zero padded)
    const float* rb;     // [N][NIp]   row bases
    const float* w;      // [NIp][NJp] alpha_i * beta_j
    float gamma;         // 2 / N^2
    int planes;
};

template <int N, int R, int P>
__global__ void __launch_bounds__(256) hfs_kernel(const HfsArgs a) {
    using D_ = HfsDims<N, R>;
    constexpr int NJp = D_::NJp, NIp = D_::NIp, NI = D_::NI, NJ = D_::NJ, XS = D_::XS, JS = D_::JS, IS = D_::IS;
    constexpr int TP = 256 / P, N4 = N / 4, HALFN = N / 2;
    static_assert(N % 4 == 0 && 256 % P == 0, "whole float4 rows, whole thread groups");
    // even-odd folding of the two large products (stages 0, 1, 5): measured 172 -> 164 us at 64 px, 308 -> 254 us at 128 px,
    // but 53 -> 57 us at 28 px (planes of 16 threads: the fold pass costs more than the shorter chains save)
    constexpr bool FOLD = (N >= 64);
    static_assert(!FOLD || R % 4 == 0, "cosine / sine column blocks of the bases must be whole float4 groups (even-odd folding)");
    // (stage 2 has only (NIp/4)*(NJp/4) register tiles per plane -- 20 at 64 px for 64 threads; splitting its K range over
    //  two adjacent lanes halves the longest serial chain but was measured SLOWER on the same box, 153.8 -> 168.7 us at
    //  4096x3x64x64: the kernel is bound by issued instructions, not by that chain, and the split adds 16 shuffles per lane)
    extern __shared__ __align__(16) float smem_hfs[];
    float* CB = smem_hfs;                       // [N][JS]
    float* RB = CB + N * JS;                    // [N][IS]
    float* Wm = RB + N * IS;                    // [NIp][NJp]
    const int p = threadIdx.x / TP, lt = threadIdx.x - p * TP;
    float* X = Wm + NIp * NJp + (size_t)p * D_::kPlane;      // [N][XS]
    float* T = X + N * XS;                      // [N][JS]
    float* V = T;                               // [N][JS]  (stage 4 onwards)
    float* Dm = T + N * JS;                     // [NIp][NJp]
    float* G = Dm + NIp * NJp;                  // [NIp][NJp]
    const int groups = (a.planes + P - 1) / P;
    // The stages of one plane only exchange data among that plane's TP threads, so they synchronise on a NAMED barrier of
    // their own (bar.sync id, TP) instead of stalling the other planes of the CTA at every stage boundary (16 % of the
    // stall samples with CTA-wide barriers, profiles/r2j_*); one warp per plane needs __syncwarp only.
    auto plane_sync = [&]() {
        if constexpr (TP == 32) __syncwarp();
        else if constexpr (TP % 32 == 0 && P <= 15) asm volatile("bar.sync %0, %1;" ::"r"(p + 1), "n"(TP) : "memory");
        else __syncthreads();
    };
    auto load_planes_async = [&](int g) {        // this thread group's plane of group g -> X (16-byte cp.async)
        const int pl = g * P + p;
        if (g < groups && pl < a.planes) {
            const float4* px = reinterpret_cast<const float4*>(a.x + (size_t)pl * N * N);
            for (int i = lt; i < N * N4; i += TP) {
                const int h = i / N4, q = i - h * N4;
                __pipeline_memcpy_async(X + h * XS + 4 * q, px + i, sizeof(float4));
            }
        }
        __pipeline_commit();
    };

    // ---- prologue: tables (all threads) and the first group's planes --------------------------------------------
    load_planes_async(blockIdx.x);
    for (int i = threadIdx.x; i < N * (NJp / 4); i += 256) {
        const int w = i / (NJp / 4), q = i - w * (NJp / 4);
        *reinterpret_cast<float4*>(CB + w * JS + 4 * q) = __ldg(reinterpret_cast<const float4*>(a.cb + w * NJp) + q);
    }
    for (int i = threadIdx.x; i < N * (NIp / 4); i += 256) {
        const int h = i / (NIp / 4), q = i - h * (NIp / 4);
        *reinterpret_cast<float4*>(RB + h * IS + 4 * q) = __ldg(reinterpret_cast<const float4*>(a.rb + h * NIp) + q);
    }
    for (int i = threadIdx.x; i < NIp * NJp / 4; i += 256)
        reinterpret_cast<float4*>(Wm)[i] = __ldg(reinterpret_cast<const float4*>(a.w) + i);

  for (int grp = blockIdx.x; grp < groups; grp += gridDim.x) {
    const int plane = grp * P + p;
    const bool live = plane < a.planes;
    __pipeline_wait_prior(0);
    if (grp == (int)blockIdx.x) __syncthreads();  // first group: the tables (loaded by all threads) are in shared memory
    else plane_sync();                            // later: only this plane's threads wrote / will read its buffers

    if constexpr (FOLD) {
    // ---- stage 0: fold every row of X into its even and odd part about w = 0 (in place; the pairs are disjoint):
    //          X[h][w] <- x[w] + x[N-w] (w = 1..N/2-1; w = 0 and N/2 are their own mirror),  X[h][N-w] <- x[N-w] - x[w].
    //      cos(k th_w) is even and sin(k th_w) odd in w, so the cosine columns of T only need the even part over
    //      w = 0..N/2 and the sine columns the odd part over w = N/2+1..N-1: half the multiply-adds of stage 1.
    if (live) {
        // items are (h, w) with w in [0, HP): HP = N/2 rounded up to a power of two keeps the index arithmetic to a shift
        // and a mask (a division by N/2 - 1 cost as much as the fold saves); consecutive lanes take consecutive w
        constexpr int HP = (HALFN <= 16) ? 16 : (HALFN <= 32 ? 32 : 64), HPS = (HP == 16) ? 4 : (HP == 32 ? 5 : 6);
        for (int e = lt; e < N * HP; e += TP) {
            const int h = e >> HPS, w = e & (HP - 1);
            if (w >= 1 && w < HALFN) {
                float* row = X + h * XS;
                const float u = row[w], v = row[N - w];
                row[w] = u + v;
                row[N - w] = v - u;
            }
        }
    }
    plane_sync();

    // ---- stage 1: T = X CB   (N x NJp): tile = rows {hg + N4*i} x columns 4jg..4jg+3; the K range is w = 0..N/2 for a
    //      tile of cosine columns and w = N/2..N-1 for sine columns (the table entry sin(k th_{N/2}) is exactly 0).  The two
    //      kinds are separate instantiations so that every loop bound and shared-memory offset is a compile-time constant
    //      (with run-time bounds the address arithmetic cost as much as the folding saved) ---------------------------------
    if (live) {
        auto tile1 = [&](auto cos_tag, const int hg, const int jg) {
            constexpr bool COS = decltype(cos_tag)::value;
            constexpr int W_LO = COS ? 0 : HALFN, W_HI = COS ? HALFN + 1 : N;
            constexpr int V_LO = (W_LO + 3) & ~3, V_HI = W_HI & ~3;          // float4-aligned body [V_LO, V_HI)
            const float* xr = X + hg * XS;
            const float* cr = CB + 4 * jg;
            float acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[i][c] = 0.0f;
            auto one = [&](const int w) {
                const float4 cv = *reinterpret_cast<const float4*>(cr + w * JS);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float xs = xr[N4 * i * XS + w];
                    acc[i][0] = fmaf(xs, cv.x, acc[i][0]);
                    acc[i][1] = fmaf(xs, cv.y, acc[i][1]);
                    acc[i][2] = fmaf(xs, cv.z, acc[i][2]);
                    acc[i][3] = fmaf(xs, cv.w, acc[i][3]);
                }
            };
#pragma unroll
            for (int w = W_LO; w < V_LO; ++w) one(w);
#pragma unroll 4
            for (int w = V_LO; w < V_HI; w += 4) {
                float4 xv[4], cv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) xv[i] = *reinterpret_cast<const float4*>(xr + N4 * i * XS + w);
#pragma unroll
                for (int q = 0; q < 4; ++q) cv[q] = *reinterpret_cast<const float4*>(cr + (w + q) * JS);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float xs[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        acc[i][0] = fmaf(xs[q], cv[q].x, acc[i][0]);
                        acc[i][1] = fmaf(xs[q], cv[q].y, acc[i][1]);
                        acc[i][2] = fmaf(xs[q], cv[q].z, acc[i][2]);
                        acc[i][3] = fmaf(xs[q], cv[q].w, acc[i][3]);
                    }
                }
            }
#pragma unroll
            for (int w = V_HI; w < W_HI; ++w) one(w);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                *reinterpret_cast<float4*>(T + (hg + N4 * i) * JS + 4 * jg) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        };
        for (int t = lt; t < N4 * (NJp / 4); t += TP) {          // cosine tiles first: at 64 px warp 0 of the plane takes them, warp 1 the sines
            const int hg = t % N4, jg = t / N4;
            if (4 * jg < R) tile1(std::true_type{}, hg, jg); else tile1(std::false_type{}, hg, jg);
        }
    }
    } else {
    // ---- stage 1: T = X CB   (N x NJp, K = N): tile = rows {hg + N4*i} x columns 4jg..4jg+3 -------------------
    if (live) {
        for (int t = lt; t < N4 * (NJp / 4); t += TP) {
            const int hg = t % N4, jg = t / N4;
            float acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[i][c] = 0.0f;
#pragma unroll 4
            for (int w4 = 0; w4 < N4; ++w4) {
                float4 xv[4], cv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) xv[i] = *reinterpret_cast<const float4*>(X + (hg + N4 * i) * XS + 4 * w4);
#pragma unroll
                for (int q = 0; q < 4; ++q) cv[q] = *reinterpret_cast<const float4*>(CB + (4 * w4 + q) * JS + 4 * jg);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float xs[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        acc[i][0] = fmaf(xs[q], cv[q].x, acc[i][0]);
                        acc[i][1] = fmaf(xs[q], cv[q].y, acc[i][1]);
                        acc[i][2] = fmaf(xs[q], cv[q].z, acc[i][2]);
                        acc[i][3] = fmaf(xs[q], cv[q].w, acc[i][3]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
                *reinterpret_cast<float4*>(T + (hg + N4 * i) * JS + 4 * jg) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
    }
    }
    plane_sync();
    load_planes_async(grp + gridDim.x);          // X is dead: fetch the next group's planes behind stages 2-5

    // ---- stage 2: D = RB^T T   (NIp x NJp, K = N): tile = 4 basis rows x 4 columns ------------------------------
    if (live) {
        for (int t = lt; t < (NIp / 4) * (NJp / 4); t += TP) {
            const int ig = t % (NIp / 4), jg = t / (NIp / 4);
            float acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[i][c] = 0.0f;
#pragma unroll 4
            for (int h = 0; h < N; ++h) {
                const float4 rv = *reinterpret_cast<const float4*>(RB + h * IS + 4 * ig);
                const float4 tv = *reinterpret_cast<const float4*>(T + h * JS + 4 * jg);
                const float rs[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    acc[i][0] = fmaf(rs[i], tv.x, acc[i][0]);
                    acc[i][1] = fmaf(rs[i], tv.y, acc[i][1]);
                    acc[i][2] = fmaf(rs[i], tv.z, acc[i][2]);
                    acc[i][3] = fmaf(rs[i], tv.w, acc[i][3]);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
                *reinterpret_cast<float4*>(Dm + (4 * ig + i) * NJp + 4 * jg) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
    }
    plane_sync();

    // ---- stage 3: G = W o D, plus the four cross terms that the unpaired frequency row -r contributes -----------
    //      basis order: rows  i = 0..R : cos(k), i = R+k : sin(k) (k = 1..R);  columns j = 0..R-1 : cos(k), j = R-1+k : sin(k)
    if (live) {
        for (int e = lt; e < NIp * NJp; e += TP) {
            const int i = e / NJp, j = e - i * NJp;
            float g = Wm[e] * Dm[e];
            if (j >= 1 && j < NJ) {
                const bool jcos = (j < R);
                const int k = jcos ? j : j - (R - 1);                    // frequency of column j (1..R-1)
                const int jc = k, js = R - 1 + k;
                if (i == 2 * R) g = jcos ? fmaf(-a.gamma, Dm[R * NJp + js], g) : fmaf(a.gamma, Dm[R * NJp + jc], g);      // row sin(R)
                if (i == R) g = jcos ? fmaf(a.gamma, Dm[2 * R * NJp + js], g) : fmaf(-a.gamma, Dm[2 * R * NJp + jc], g);   // row cos(R)
            }
            G[e] = (i < NI && j < NJ) ? g : 0.0f;
        }
    }
    plane_sync();

    // ---- stage 4: V = RB G   (N x NJp, K = NI): tile = rows {hg + N4*i} x columns 4jg..4jg+3 --------------------
    if (live) {
        for (int t = lt; t < N4 * (NJp / 4); t += TP) {
            const int hg = t % N4, jg = t / N4;
            float acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[i][c] = 0.0f;
#pragma unroll
            for (int k = 0; k < NI; ++k) {
                const float4 gv = *reinterpret_cast<const float4*>(G + k * NJp + 4 * jg);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float rv = RB[(hg + N4 * i) * IS + k];
                    acc[i][0] = fmaf(rv, gv.x, acc[i][0]);
                    acc[i][1] = fmaf(rv, gv.y, acc[i][1]);
                    acc[i][2] = fmaf(rv, gv.z, acc[i][2]);
                    acc[i][3] = fmaf(rv, gv.w, acc[i][3]);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
                *reinterpret_cast<float4*>(V + (hg + N4 * i) * JS + 4 * jg) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
    }
    plane_sync();

    if constexpr (FOLD) {
    // ---- stage 5: y = V CB^T   (N x N): with Ye = (cosine columns of V) . cos(k th_w) and Yo = (sine columns) . sin(k th_w),
    //      y[h][w] = Ye + Yo and y[h][N-w] = Ye - Yo: one tile = rows {hg + N4*i} x columns {wg + WS*c} of the HALF plane
    //      w < N/2 with an even and an odd accumulator each (half the multiply-adds); the column w = N/2 (Yo == 0) is a
    //      short pass of its own --------------------------------------------------------------------------------------
    if (live) {
        float* py = a.y + (size_t)plane * N * N;
        const float* pa = a.add ? a.add + (size_t)plane * N * N : nullptr;
        constexpr int WS = (HALFN + 3) / 4;
        constexpr bool EXACT = (HALFN % 4 == 0);                  // every column slot of a tile is a real column
        for (int t = lt; t < N4 * WS; t += TP) {
            const int wg = t % WS, hg = t / WS;
            float ae[4][4], ao[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c < 4; ++c) { ae[i][c] = 0.0f; ao[i][c] = 0.0f; }
#pragma unroll
            for (int q = 0; q < NJp / 4; ++q) {
                float4 vv[4], cv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) vv[i] = *reinterpret_cast<const float4*>(V + (hg + N4 * i) * JS + 4 * q);
#pragma unroll
                for (int c = 0; c < 4; ++c) cv[c] = *reinterpret_cast<const float4*>(CB + (EXACT ? wg + WS * c : min(wg + WS * c, HALFN)) * JS + 4 * q);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float s = (4 * q < R) ? ae[i][c] : ao[i][c];
                        s = fmaf(vv[i].x, cv[c].x, s);
                        s = fmaf(vv[i].y, cv[c].y, s);
                        s = fmaf(vv[i].z, cv[c].z, s);
                        s = fmaf(vv[i].w, cv[c].w, s);
                        if (4 * q < R) ae[i][c] = s; else ao[i][c] = s;
                    }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int w = wg + WS * c;
                    if (EXACT || w < HALFN) {
                        const int o1 = (hg + N4 * i) * N + w;
                        const float y1 = ae[i][c] + ao[i][c];
                        __stcs(py + o1, pa ? y1 + __ldcs(pa + o1) : y1);
                        if (w != 0) {
                            const int o2 = (hg + N4 * i) * N + N - w;
                            const float y2 = ae[i][c] - ao[i][c];
                            __stcs(py + o2, pa ? y2 + __ldcs(pa + o2) : y2);
                        }
                    }
                }
        }
        for (int h = lt; h < N; h += TP) {                      // the self-mirrored column w = N/2
            float se = 0.0f, so = 0.0f;
#pragma unroll
            for (int j = 0; j < NJp; ++j) {
                const float p_ = V[h * JS + j], c_ = CB[HALFN * JS + j];
                if (j < R) se = fmaf(p_, c_, se); else so = fmaf(p_, c_, so);
            }
            const int o = h * N + HALFN;
            const float y1 = se + so;
            __stcs(py + o, pa ? y1 + __ldcs(pa + o) : y1);
        }
    }
    } else {
    // ---- stage 5: y = V CB^T   (N x N, K = NJp): tile = rows {hg + N4*i} x columns {wg + N4*c} -------------------
    if (live) {
        float* py = a.y + (size_t)plane * N * N;
        const float* pa = a.add ? a.add + (size_t)plane * N * N : nullptr;
        for (int t = lt; t < N4 * N4; t += TP) {
            const int wg = t % N4, hg = t / N4;
            float acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[i][c] = 0.0f;
#pragma unroll
            for (int q = 0; q < NJp / 4; ++q) {
                float4 vv[4], cv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) vv[i] = *reinterpret_cast<const float4*>(V + (hg + N4 * i) * JS + 4 * q);
#pragma unroll
                for (int c = 0; c < 4; ++c) cv[c] = *reinterpret_cast<const float4*>(CB + (wg + N4 * c) * JS + 4 * q);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float s = acc[i][c];
                        s = fmaf(vv[i].x, cv[c].x, s);
                        s = fmaf(vv[i].y, cv[c].y, s);
                        s = fmaf(vv[i].z, cv[c].z, s);
                        s = fmaf(vv[i].w, cv[c].w, s);
                        acc[i][c] = s;
                    }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                        const int o = (hg + N4 * i) * N + wg + N4 * c;
                        __stcs(py + o, pa ? acc[i][c] + __ldcs(pa + o) : acc[i][c]);
                    }
        }
    }
    }
  }   // persistent loop over plane groups (the barrier at its top also protects V against the next stage 1)
}

}  // namespace ee

namespace ee {

// -------------------------------------------------------------------------------------------------------------------
// Large planes (the fast-AT schedule of ImageNet/fgsm_imagenet: 128 px / r 12, 224 px / r 16, 288 px / r 18): the plane does not fit
// in shared memory next to the tables.  One plane per (persistent) CTA;
// x is STREAMED through two cp.async row-block buffers of RBK = 16 (8 at 288 px, 32 at 128 px) rows while T (N x NJp) is accumulated block by block;
// D, G, V live in shared memory as above; y is produced row block by row block straight into global memory.
// A 16-row block only offers 4 x NJp/4 register tiles, so the K range (w) of stage 1 is split over the 8 lanes that share
// a tile and reduced with an xor-butterfly ((p0+p1)+(p2+p3))+((p4+p5)+(p6+p7)); stage 2 splits its K range (h) over 2
// lanes.  The oracle evaluates the same partial chains and the same trees (ks1 = 8, ks2 = 2).
// -------------------------------------------------------------------------------------------------------------------
template <int N, int R>
struct HfsRowsDims {
    using D_ = HfsDims<N, R>;
    // 224 px: HALF planes (112 rows, 102 KB) fit next to the tables, so stage 1 runs like the whole-plane kernel (one 4 x 4
    // tile per thread, no lane split) on a single buffer; the second half's load is exposed (~1.5 us per plane), the next
    // plane's first half streams in behind stages 2-5.  288 px: 8-row blocks, double buffered, K range split over lanes.
    static constexpr bool HALF = (N == 224);
    static constexpr int RBK = HALF ? N / 2 : ((N >= 256) ? 8 : 16), KS2 = 2, NBUF = HALF ? 1 : 2;
    static constexpr int KS1 = HALF ? 1 : (((RBK / 4) * (D_::NJp / 4) * 8 <= 256) ? 8 : 4);     // lanes sharing a stage-1 tile
    static constexpr int kFloats = N * D_::JS + N * D_::IS + D_::NIp * D_::NJp     // CB, RB, W
                                   + N * D_::JS + 2 * D_::NIp * D_::NJp           // T (= V), D, G
                                   + NBUF * RBK * D_::XS                          // x row-block buffer(s)
                                   + (KS1 > 1 ? D_::NJp * D_::XS : 0);            // CB transposed (lane-split stage 1)
};

template <int N, int R>
__global__ void __launch_bounds__(256, 1) hfs_rows_kernel(const HfsArgs a) {
    using D_ = HfsDims<N, R>;
    using Q_ = HfsRowsDims<N, R>;
    constexpr int NJp = D_::NJp, NIp = D_::NIp, NI = D_::NI, NJ = D_::NJ, XS = D_::XS, JS = D_::JS, IS = D_::IS;
    constexpr int RBK = Q_::RBK, KS1 = Q_::KS1, KS2 = Q_::KS2, NBUF = Q_::NBUF, N4 = N / 4, NBLK = N / RBK, HALFN = N / 2;
    static_assert(N % RBK == 0 && N % (4 * KS1) == 0 && N % KS2 == 0, "row blocks and K splits divide the plane");
    static_assert((RBK / 4) * (NJp / 4) * KS1 <= 256, "stage 1 fits one pass of the CTA");
    static_assert(NBUF == 1 || NBLK % 2 == 0, "double buffering: the next plane's first block lands in buffer 0");
    extern __shared__ __align__(16) float smem_hfs[];
    float* CB = smem_hfs;
    float* RB = CB + N * JS;
    float* Wm = RB + N * IS;
    float* T = Wm + NIp * NJp;                 // [N][JS], later V
    float* V = T;
    float* Dm = T + N * JS;
    float* G = Dm + NIp * NJp;
    float* XB = G + NIp * NJp;                 // [2][RBK][XS]
    // CB transposed, [NJp][XS]: the lanes that split stage 1's K range read x AND the basis at offsets of N/KS1 floats
    // along w (distinct bank groups); with the [w][j] layout their rows would sit 16 or 0 banks apart (4-way conflicts:
    // 43 % of all shared-memory wavefronts in the first version, profiles/r1g_ncu_full_hfs224_first.txt)
    float* CBt = XB + NBUF * RBK * XS;
    const int tid = threadIdx.x;

    auto load_block_async = [&](int plane, int blk, int buf) {       // rows [blk*RBK, +RBK) of `plane` -> XB[buf]
        if (plane < a.planes) {
            const float4* px = reinterpret_cast<const float4*>(a.x + (size_t)plane * N * N + (size_t)blk * RBK * N);
            float* dst = XB + buf * RBK * XS;
            for (int i = tid; i < RBK * N4; i += 256) {
                const int h = i / N4, q = i - h * N4;
                __pipeline_memcpy_async(dst + h * XS + 4 * q, px + i, sizeof(float4));
            }
        }
        __pipeline_commit();
    };

    load_block_async(blockIdx.x, 0, 0);
    for (int i = tid; i < N * (NJp / 4); i += 256) {
        const int w = i / (NJp / 4), q = i - w * (NJp / 4);
        *reinterpret_cast<float4*>(CB + w * JS + 4 * q) = __ldg(reinterpret_cast<const float4*>(a.cb + w * NJp) + q);
    }
    for (int i = tid; i < N * (NIp / 4); i += 256) {
        const int h = i / (NIp / 4), q = i - h * (NIp / 4);
        *reinterpret_cast<float4*>(RB + h * IS + 4 * q) = __ldg(reinterpret_cast<const float4*>(a.rb + h * NIp) + q);
    }
    for (int i = tid; i < NIp * NJp / 4; i += 256) reinterpret_cast<float4*>(Wm)[i] = __ldg(reinterpret_cast<const float4*>(a.w) + i);
    if (KS1 > 1) {
        for (int i = tid; i < N * NJp; i += 256) {
            const int w = i / NJp, j = i - w * NJp;
            CBt[j * XS + w] = __ldg(a.cb + i);
        }
    }

    for (int plane = blockIdx.x; plane < a.planes; plane += gridDim.x) {
        // ---- stage 1: T = x CB, row block by row block; block b+1 (or block 0 of the next plane) streams in meanwhile ----
        for (int blk = 0; blk < NBLK; ++blk) {
            __pipeline_wait_prior(0);
            __syncthreads();                   // block `blk` is in its buffer; with two buffers the other one is free again
            constexpr int kNextBuf = 0;
            (void)kNextBuf;
            if (NBUF == 2) {
                if (blk + 1 < NBLK) load_block_async(plane, blk + 1, (blk + 1) & 1);
                else load_block_async(plane + gridDim.x, 0, (blk + 1) & 1);
            }
            const float* X = XB + (NBUF == 2 ? (blk & 1) : 0) * RBK * XS;
            if constexpr (KS1 == 1) {
                // Even / odd folding as in hfs_kernel: fold every row of the block in place (x[w] + x[N-w] for w < N/2,
                // x[w] - x[N-w] for w > N/2), then one 4 x 4 tile per thread over HALF the K range -- w = 0..N/2 for a tile
                // of cosine columns, w = N/2..N-1 for sine columns (table entry exactly 0 at N/2)
                {
                    constexpr int HP = 128;                                  // N/2 = 112 rounded up to a power of two
                    static_assert(HALFN <= HP && R % 4 == 0 && HALFN % 4 == 0, "fold indexing");
                    float* Xw = XB;
                    for (int e = tid; e < RBK * HP; e += 256) {
                        const int h = e >> 7, w = e & (HP - 1);
                        if (w >= 1 && w < HALFN) {
                            float* row = Xw + h * XS;
                            const float u = row[w], v = row[N - w];
                            row[w] = u + v;
                            row[N - w] = v - u;
                        }
                    }
                }
                __syncthreads();
                auto tile1 = [&](auto cos_tag, const int hg, const int jg) {
                    constexpr bool COS = decltype(cos_tag)::value;
                    constexpr int W_LO = COS ? 0 : HALFN, W_HI = COS ? HALFN + 1 : N, V_HI = W_HI & ~3;
                    const float* xr = X + hg * XS;
                    const float* cr = CB + 4 * jg;
                    float acc[4][4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int c = 0; c < 4; ++c) acc[i][c] = 0.0f;
#pragma unroll 4
                    for (int w = W_LO; w < V_HI; w += 4) {
                        float4 xv[4], cv[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) xv[i] = *reinterpret_cast<const float4*>(xr + (RBK / 4) * i * XS + w);
#pragma unroll
                        for (int q = 0; q < 4; ++q) cv[q] = *reinterpret_cast<const float4*>(cr + (w + q) * JS);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float xs[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                acc[i][0] = fmaf(xs[q], cv[q].x, acc[i][0]);
                                acc[i][1] = fmaf(xs[q], cv[q].y, acc[i][1]);
                                acc[i][2] = fmaf(xs[q], cv[q].z, acc[i][2]);
                                acc[i][3] = fmaf(xs[q], cv[q].w, acc[i][3]);
                            }
                        }
                    }
#pragma unroll
                    for (int w = V_HI; w < W_HI; ++w) {                  // the cosine tiles' last term, w = N/2
                        const float4 cv = *reinterpret_cast<const float4*>(cr + w * JS);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float xs = xr[(RBK / 4) * i * XS + w];
                            acc[i][0] = fmaf(xs, cv.x, acc[i][0]);
                            acc[i][1] = fmaf(xs, cv.y, acc[i][1]);
                            acc[i][2] = fmaf(xs, cv.z, acc[i][2]);
                            acc[i][3] = fmaf(xs, cv.w, acc[i][3]);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        *reinterpret_cast<float4*>(T + (blk * RBK + hg + (RBK / 4) * i) * JS + 4 * jg) =
                            make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
                };
                for (int t = tid; t < (RBK / 4) * (NJp / 4); t += 256) {
                    const int hg = t % (RBK / 4), jg = t / (RBK / 4);
                    if (4 * jg < R) tile1(std::true_type{}, hg, jg); else tile1(std::false_type{}, hg, jg);
                }
                __syncthreads();               // every thread is done with the single buffer: refill it
                if (blk + 1 < NBLK) load_block_async(plane, blk + 1, 0);
                else load_block_async(plane + gridDim.x, 0, 0);
                continue;
            }
            const int ks = tid % KS1;
            const bool tile_ok = (tid / KS1) < (RBK / 4) * (NJp / 4);      // lanes without a tile still take part in the shuffles
            const int tile = tile_ok ? tid / KS1 : 0;
            {
                const int hg = tile % (RBK / 4), jg = tile / (RBK / 4);
                float acc[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[i][c] = 0.0f;
                constexpr int W4S = N4 / KS1;                      // float4 steps per K split
#pragma unroll
                for (int s = 0; s < W4S; ++s) {
                    const int w4 = ks * W4S + s;
                    float4 xv[4], ct[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) xv[i] = *reinterpret_cast<const float4*>(X + (hg + (RBK / 4) * i) * XS + 4 * w4);
#pragma unroll
                    for (int c = 0; c < 4; ++c) ct[c] = *reinterpret_cast<const float4*>(CBt + (4 * jg + c) * XS + 4 * w4);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float xs[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const float cs[4] = {ct[c].x, ct[c].y, ct[c].z, ct[c].w};
#pragma unroll
                            for (int q = 0; q < 4; ++q) acc[i][c] = fmaf(xs[q], cs[q], acc[i][c]);
                        }
                    }
                }
#pragma unroll
                for (int st = 1; st < KS1; st <<= 1)
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int c = 0; c < 4; ++c) acc[i][c] = acc[i][c] + __shfl_xor_sync(0xffffffffu, acc[i][c], st);
                if (ks == 0 && tile_ok) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        *reinterpret_cast<float4*>(T + (blk * RBK + hg + (RBK / 4) * i) * JS + 4 * jg) =
                            make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
                }
            }
        }
        __syncthreads();

        // ---- stage 2: D = RB^T T, K (= h) split over two adjacent lanes --------------------------------------------
        {
            const int ks = tid % KS2;
            const bool t_ok = (tid / KS2) < (NIp / 4) * (NJp / 4);
            const int t = t_ok ? tid / KS2 : 0;
            {
                const int ig = t % (NIp / 4), jg = t / (NIp / 4);
                float acc[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[i][c] = 0.0f;
#pragma unroll 4
                for (int hh = 0; hh < N / KS2; ++hh) {
                    const int h = ks * (N / KS2) + hh;
                    const float4 rv = *reinterpret_cast<const float4*>(RB + h * IS + 4 * ig);
                    const float4 tv = *reinterpret_cast<const float4*>(T + h * JS + 4 * jg);
                    const float rs[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        acc[i][0] = fmaf(rs[i], tv.x, acc[i][0]);
                        acc[i][1] = fmaf(rs[i], tv.y, acc[i][1]);
                        acc[i][2] = fmaf(rs[i], tv.z, acc[i][2]);
                        acc[i][3] = fmaf(rs[i], tv.w, acc[i][3]);
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[i][c] = acc[i][c] + __shfl_xor_sync(0xffffffffu, acc[i][c], 1);
                if (ks == 0 && t_ok) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        *reinterpret_cast<float4*>(Dm + (4 * ig + i) * NJp + 4 * jg) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
                }
            }
        }
        __syncthreads();

        // ---- stage 3: G = W o D + cross terms of the frequency row -r (as in hfs_kernel) ---------------------------
        for (int e = tid; e < NIp * NJp; e += 256) {
            const int i = e / NJp, j = e - i * NJp;
            float g = Wm[e] * Dm[e];
            if (j >= 1 && j < NJ) {
                const bool jcos = (j < R);
                const int k = jcos ? j : j - (R - 1);
                const int jc = k, js = R - 1 + k;
                if (i == 2 * R) g = jcos ? fmaf(-a.gamma, Dm[R * NJp + js], g) : fmaf(a.gamma, Dm[R * NJp + jc], g);
                if (i == R) g = jcos ? fmaf(a.gamma, Dm[2 * R * NJp + js], g) : fmaf(-a.gamma, Dm[2 * R * NJp + jc], g);
            }
            G[e] = (i < NI && j < NJ) ? g : 0.0f;
        }
        __syncthreads();

        // ---- stage 4: V = RB G (overwrites T) -----------------------------------------------------------------------
        for (int t = tid; t < N4 * (NJp / 4); t += 256) {
            const int hg = t % N4, jg = t / N4;
            float acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[i][c] = 0.0f;
#pragma unroll
            for (int k = 0; k < NI; ++k) {
                const float4 gv = *reinterpret_cast<const float4*>(G + k * NJp + 4 * jg);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float rv = RB[(hg + N4 * i) * IS + k];
                    acc[i][0] = fmaf(rv, gv.x, acc[i][0]);
                    acc[i][1] = fmaf(rv, gv.y, acc[i][1]);
                    acc[i][2] = fmaf(rv, gv.z, acc[i][2]);
                    acc[i][3] = fmaf(rv, gv.w, acc[i][3]);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
                *reinterpret_cast<float4*>(V + (hg + N4 * i) * JS + 4 * jg) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
        __syncthreads();

        // ---- stage 5: y = V CB^T, tiles of rows {hg + N4*i} x columns {wg + N4*c}; with the folded kernel (KS1 == 1) a
        //      tile covers columns of the half plane w < N/2 with an even and an odd accumulator: y[w] = Ye + Yo, y[N-w] = Ye - Yo
        if constexpr (KS1 == 1) {
            float* py = a.y + (size_t)plane * N * N;
            const float* pa = a.add ? a.add + (size_t)plane * N * N : nullptr;
            constexpr int WS = HALFN / 4;
            for (int t = tid; t < N4 * WS; t += 256) {
                const int wg = t % WS, hg = t / WS;
                float ae[4][4], ao[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int c = 0; c < 4; ++c) { ae[i][c] = 0.0f; ao[i][c] = 0.0f; }
#pragma unroll
                for (int q = 0; q < NJp / 4; ++q) {
                    float4 vv[4], cv[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) vv[i] = *reinterpret_cast<const float4*>(V + (hg + N4 * i) * JS + 4 * q);
#pragma unroll
                    for (int c = 0; c < 4; ++c) cv[c] = *reinterpret_cast<const float4*>(CB + (wg + WS * c) * JS + 4 * q);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            float sacc = (4 * q < R) ? ae[i][c] : ao[i][c];
                            sacc = fmaf(vv[i].x, cv[c].x, sacc);
                            sacc = fmaf(vv[i].y, cv[c].y, sacc);
                            sacc = fmaf(vv[i].z, cv[c].z, sacc);
                            sacc = fmaf(vv[i].w, cv[c].w, sacc);
                            if (4 * q < R) ae[i][c] = sacc; else ao[i][c] = sacc;
                        }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int w = wg + WS * c;
                        const int o1 = (hg + N4 * i) * N + w;
                        const float y1 = ae[i][c] + ao[i][c];
                        __stcs(py + o1, pa ? y1 + __ldcs(pa + o1) : y1);
                        if (w != 0) {
                            const int o2 = (hg + N4 * i) * N + N - w;
                            const float y2 = ae[i][c] - ao[i][c];
                            __stcs(py + o2, pa ? y2 + __ldcs(pa + o2) : y2);
                        }
                    }
            }
            for (int h = tid; h < N; h += 256) {                      // the self-mirrored column w = N/2
                float se = 0.0f, so = 0.0f;
#pragma unroll
                for (int j = 0; j < NJp; ++j) {
                    const float p_ = V[h * JS + j], c_ = CB[HALFN * JS + j];
                    if (j < R) se = fmaf(p_, c_, se); else so = fmaf(p_, c_, so);
                }
                const int o = h * N + HALFN;
                const float y1 = se + so;
                __stcs(py + o, pa ? y1 + __ldcs(pa + o) : y1);
            }
        } else {
            float* py = a.y + (size_t)plane * N * N;
            const float* pa = a.add ? a.add + (size_t)plane * N * N : nullptr;
            for (int t = tid; t < N4 * N4; t += 256) {
                const int wg = t % N4, hg = t / N4;
                float acc[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[i][c] = 0.0f;
#pragma unroll
                for (int q = 0; q < NJp / 4; ++q) {
                    float4 vv[4], cv[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) vv[i] = *reinterpret_cast<const float4*>(V + (hg + N4 * i) * JS + 4 * q);
#pragma unroll
                    for (int c = 0; c < 4; ++c) cv[c] = *reinterpret_cast<const float4*>(CB + (wg + N4 * c) * JS + 4 * q);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            float s = acc[i][c];
                            s = fmaf(vv[i].x, cv[c].x, s);
                            s = fmaf(vv[i].y, cv[c].y, s);
                            s = fmaf(vv[i].z, cv[c].z, s);
                            s = fmaf(vv[i].w, cv[c].w, s);
                            acc[i][c] = s;
                        }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int o = (hg + N4 * i) * N + wg + N4 * c;
                        __stcs(py + o, pa ? acc[i][c] + __ldcs(pa + o) : acc[i][c]);
                    }
            }
        }
        // the barrier at the top of the next plane's first row block protects V (= T) and the x buffers
    }
}

}  // namespace ee
