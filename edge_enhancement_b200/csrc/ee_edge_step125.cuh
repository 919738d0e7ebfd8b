// ee_edge_step125.cuh -- fused kernels for CannyFilter_step125_1 (+ blend), forward and adjoint.
//
// Replaces (reference paths): utils/core.py:549-585 (filter forward), the blend
// Tiny_ImageNet/models_tinyimagenet/resnet_EE.py:189-191, and the autograd graph through both.
//
// Work decomposition: one CTA owns a full-width strip of TH rows of one image.  The stencil
// pipeline runs on single-channel planes in shared memory because the channel sum commutes with
// the blur / replicate-pad / Sobel chain (DESIGN.md "Channel-sum first"):
//
//   forward :  S = sum_c x_c  (rows r0-2..r1+2)  ->  Bl = blur(S) (r0-1..r1+1)  ->  edge (r0..r1)
//              -> out_c = clamp(base_c + w*edge)            HBM: read x, base ; write out
//   backward:  S (r0-4..r1+4) -> Bl (r0-3..r1+3) -> A,Bv = dL/dSg{x,y} (r0-2..r1+2; needs g_out,
//              base there) -> GB = dL/dblur (r0-1..r1+1) -> g_s (r0..r1) -> g_x_c = g_s
//              HBM: read g_out, x, base ; write g_x, g_base
//
// Row halos are re-read by the neighbouring strip's CTA (an L2 hit, not an HBM read).  Image
// borders are handled by clamping (forward, replicate) and by folding the padded ring back onto
// the border (adjoint) exactly as the oracle does.
#pragma once
#include "ee_device.cuh"

namespace ee {

struct EEStride3 { int64_t b, c, h; };

struct EdgeArgs {
    const float* x;       // [B,C,H,W]
    const float* base;    // [B,C,H,W]   (blend only)
    const float* g_in;    // bwd: g_out [B,C,H,W] (blend) or g_edge [B,1,H,W]
    float* out;           // fwd blend: [B,C,H,W]
    float* edge;          // fwd: [B,1,H,W] or null
    float* g_x;           // bwd: [B,C,H,W] or null
    float* g_base;        // bwd blend: [B,C,H,W] or null
    int B, C, H, W;
    int TH, tiles_per_img;
    int GX, RY;           // thread t -> (tx = t % GX, ty = t / GX); rows advance by RY
    float c0, c1, c2, fC, alpha, low, high, w;
    int variant, has_low, has_high, hyst;
    int nan_compat;       // EE_FLAG_NAN_COMPAT: dL/dSgx = dL/dSgy = NaN where the magnitude is exactly 0 (generic backward kernels)
    // element strides (batch, channel, row) of the image tensors; the column stride is always 1.  Dense NCHW is
    // {C*H*W, H*W, W}; the shape-generic kernels honour them (sliced batches, channel slices, crops: ee_*_strided_f32),
    // the tuned kernels require dense tensors.
    EEStride3 sx, sbase, sg, sout, sgx, sgbase, sedge;
    int strided;          // some tensor is not dense: shape-generic kernels only
};

// element offset of (image b, channel c, row r, column col)
__device__ __forceinline__ size_t at(const EEStride3& s, int b, int c, int r, int col) {
    return (size_t)((int64_t)b * s.b + (int64_t)c * s.c + (int64_t)r * s.h) + col;
}

#define EE_FOR_TILE(row_lo, row_hi)                                     \
    if (ty < a.RY)                                                      \
        for (int row = (row_lo) + ty; row < (row_hi); row += a.RY)      \
            for (int g = tx; g < G; g += a.GX)

// S plane: channel sum of x rows [lo, hi) of image b
template <int VEC, int NC>
__device__ __forceinline__ void stage_channel_sum(const EdgeArgs& a, const float* __restrict__ xb, float* S,
                                                  int lo, int hi, int G, int tx, int ty) {
    const int C = NC ? NC : a.C;
    const int W = a.W;
    const size_t hw = (size_t)a.sx.c;                  // channel stride of x
    EE_FOR_TILE(lo, hi) {
        const int col = g * VEC;
        const float* px = xb + (size_t)row * (size_t)a.sx.h + col;
        float acc[VEC];
        ldg_vec<VEC>(px, acc);
        if (NC == 3) {
            float t1[VEC], t2[VEC];
            ldg_vec<VEC>(px + hw, t1);
            ldg_vec<VEC>(px + 2 * hw, t2);
#pragma unroll
            for (int k = 0; k < VEC; ++k) acc[k] = (acc[k] + t1[k]) + t2[k];
        } else {
            for (int c = 1; c < C; ++c) {
                float t[VEC];
                ldg_vec<VEC>(px + (size_t)c * hw, t);
#pragma unroll
                for (int k = 0; k < VEC; ++k) acc[k] = acc[k] + t[k];
            }
        }
        st_vec<VEC>(S + (size_t)(row - lo) * W + col, acc);
    }
}

// Bl plane rows [lo, hi) from S plane (rows [s_lo, ..)), replicate padding
template <int VEC>
__device__ __forceinline__ void stage_blur(const EdgeArgs& a, const float* S, int s_lo, float* Bl, int lo,
                                           int hi, int G, int tx, int ty) {
    const int W = a.W, H = a.H;
    EE_FOR_TILE(lo, hi) {
        const int col = g * VEC;
        const int ru = max(row - 1, 0), rd = min(row + 1, H - 1);
        float u[VEC + 2], m[VEC + 2], d[VEC + 2], o[VEC];
        load_win<VEC, false>(S + (size_t)(ru - s_lo) * W, col, W, u);
        load_win<VEC, false>(S + (size_t)(row - s_lo) * W, col, W, m);
        load_win<VEC, false>(S + (size_t)(rd - s_lo) * W, col, W, d);
        gauss3<VEC>(u, m, d, a.c0, a.c1, a.c2, o);
        st_vec<VEC>(Bl + (size_t)(row - lo) * W + col, o);
    }
}

// gx1, gy1 of VEC pixels at (row, col) from the Bl plane (replicate padding)
template <int VEC>
__device__ __forceinline__ void sobel_at(const EdgeArgs& a, const float* Bl, int b_lo, int row, int col,
                                         float (&gx1)[VEC], float (&gy1)[VEC]) {
    const int W = a.W, H = a.H;
    const int ru = max(row - 1, 0), rd = min(row + 1, H - 1);
    float u[VEC + 2], m[VEC + 2], d[VEC + 2];
    load_win<VEC, false>(Bl + (size_t)(ru - b_lo) * W, col, W, u);
    load_win<VEC, false>(Bl + (size_t)(row - b_lo) * W, col, W, m);
    load_win<VEC, false>(Bl + (size_t)(rd - b_lo) * W, col, W, d);
    sobel3<VEC>(u, m, d, a.fC, gx1, gy1);
}

// -------------------------------------------------------------------------------------------
// forward
// -------------------------------------------------------------------------------------------
template <int VEC, int NC, bool BLEND>
__global__ void __launch_bounds__(256) edge_fwd_step125_kernel(const EdgeArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int b = blockIdx.x / a.tiles_per_img;
    const int ti = blockIdx.x - b * a.tiles_per_img;
    const int H = a.H, W = a.W;
    const int C = NC ? NC : a.C;
    const int r0 = ti * a.TH, r1 = min(r0 + a.TH, H);
    const int G = (W + VEC - 1) / VEC;
    const int tx = threadIdx.x % a.GX, ty = threadIdx.x / a.GX;
    const size_t hw = (size_t)H * W;

    const int s_lo = max(r0 - 2, 0), s_hi = min(r1 + 2, H);
    const int b_lo = max(r0 - 1, 0), b_hi = min(r1 + 1, H);
    float* S = smem;
    float* Bl = smem + (size_t)(a.TH + 4) * W;

    stage_channel_sum<VEC, NC>(a, a.x + (size_t)((int64_t)b * a.sx.b), S, s_lo, s_hi, G, tx, ty);
    __syncthreads();
    stage_blur<VEC>(a, S, s_lo, Bl, b_lo, b_hi, G, tx, ty);
    __syncthreads();

    EE_FOR_TILE(r0, r1) {
        const int col = g * VEC;
        const size_t pix = (size_t)row * W + col;
        float bs[NC ? NC : 1][VEC];
        if (BLEND && NC) {
#pragma unroll
            for (int c = 0; c < NC; ++c) ldg_vec<VEC>(a.base + at(a.sbase, b, c, row, col), bs[c]);
        }
        float gx1[VEC], gy1[VEC], e[VEC];
        sobel_at<VEC>(a, Bl, b_lo, row, col, gx1, gy1);
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            const float mag = magnitude(gx1[k], gy1[k]);
            const float magm = (mag < a.alpha) ? 0.0f : mag;       // core.py:574-575
            e[k] = to_compare(magm, a.high);                      // core.py:578-583
        }
        if (a.edge) stg_vec<VEC>(a.edge + at(a.sedge, b, 0, row, col), e);
        if (BLEND) {
            float we[VEC];
#pragma unroll
            for (int k = 0; k < VEC; ++k) we[k] = a.w * e[k];
            if (NC) {
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    float o[VEC];
#pragma unroll
                    for (int k = 0; k < VEC; ++k) o[k] = clamp01_nan(bs[c][k] + we[k]);
                    stg_vec<VEC>(a.out + at(a.sout, b, c, row, col), o);
                }
            } else {
                for (int c = 0; c < C; ++c) {
                    float t[VEC], o[VEC];
                    ldg_vec<VEC>(a.base + at(a.sbase, b, c, row, col), t);
#pragma unroll
                    for (int k = 0; k < VEC; ++k) o[k] = clamp01_nan(t[k] + we[k]);
                    stg_vec<VEC>(a.out + at(a.sout, b, c, row, col), o);
                }
            }
        }
    }
}

// -------------------------------------------------------------------------------------------
// adjoint helpers: column-folded rows of the two transposed stencils
// -------------------------------------------------------------------------------------------
// plane row pointer or nullptr when the row is outside the image (zero extension)
__device__ __forceinline__ const float* zrow(const float* P, int lo, int r, int H, int W) {
    return (r >= 0 && r < H) ? P + (size_t)(r - lo) * W : nullptr;
}

// F(p, col..col+VEC) of the Sobel adjoint: T(p,q) with the ring columns -1 / W folded onto 0 / W-1
template <int VEC>
__device__ __forceinline__ void sobel_adj_row(const float* A, const float* Bv, int lo, int p, int col, int H,
                                              int W, float (&o)[VEC]) {
    const float* au = zrow(A, lo, p - 1, H, W); const float* am = zrow(A, lo, p, H, W); const float* ad = zrow(A, lo, p + 1, H, W);
    const float* bu = zrow(Bv, lo, p - 1, H, W); const float* bm = zrow(Bv, lo, p, H, W); const float* bd = zrow(Bv, lo, p + 1, H, W);
    float wau[VEC + 2], wam[VEC + 2], wad[VEC + 2], wbu[VEC + 2], wbm[VEC + 2], wbd[VEC + 2];
    load_win<VEC, true>(au, col, W, wau); load_win<VEC, true>(am, col, W, wam); load_win<VEC, true>(ad, col, W, wad);
    load_win<VEC, true>(bu, col, W, wbu); load_win<VEC, true>(bm, col, W, wbm); load_win<VEC, true>(bd, col, W, wbd);
    sobel3_adj<VEC>(wau, wam, wad, wbu, wbm, wbd, o);
    if (col == 0 || col + VEC >= W) {
        float sau[3], sam[3], sad[3], sbu[3], sbm[3], sbd[3], t[1];
        if (col == 0) {
            load_win1_zero(au, -1, W, sau); load_win1_zero(am, -1, W, sam); load_win1_zero(ad, -1, W, sad);
            load_win1_zero(bu, -1, W, sbu); load_win1_zero(bm, -1, W, sbm); load_win1_zero(bd, -1, W, sbd);
            sobel3_adj<1>(sau, sam, sad, sbu, sbm, sbd, t);
            o[0] = o[0] + t[0];
        }
        if (col + VEC >= W) {
            load_win1_zero(au, W, W, sau); load_win1_zero(am, W, W, sam); load_win1_zero(ad, W, W, sad);
            load_win1_zero(bu, W, W, sbu); load_win1_zero(bm, W, W, sbm); load_win1_zero(bd, W, W, sbd);
            sobel3_adj<1>(sau, sam, sad, sbu, sbm, sbd, t);
            o[W - 1 - col] = o[W - 1 - col] + t[0];
        }
    }
}

template <int VEC>
__device__ __forceinline__ void gauss_adj_row(const EdgeArgs& a, const float* GB, int lo, int p, int col, int H,
                                              int W, float (&o)[VEC]) {
    const float* gu = zrow(GB, lo, p - 1, H, W); const float* gm = zrow(GB, lo, p, H, W); const float* gd = zrow(GB, lo, p + 1, H, W);
    float wu[VEC + 2], wm[VEC + 2], wd[VEC + 2];
    load_win<VEC, true>(gu, col, W, wu); load_win<VEC, true>(gm, col, W, wm); load_win<VEC, true>(gd, col, W, wd);
    gauss3<VEC>(wu, wm, wd, a.c0, a.c1, a.c2, o);
    if (col == 0 || col + VEC >= W) {
        float su[3], sm[3], sd[3], t[1];
        if (col == 0) {
            load_win1_zero(gu, -1, W, su); load_win1_zero(gm, -1, W, sm); load_win1_zero(gd, -1, W, sd);
            gauss3<1>(su, sm, sd, a.c0, a.c1, a.c2, t);
            o[0] = o[0] + t[0];
        }
        if (col + VEC >= W) {
            load_win1_zero(gu, W, W, su); load_win1_zero(gm, W, W, sm); load_win1_zero(gd, W, W, sd);
            gauss3<1>(su, sm, sd, a.c0, a.c1, a.c2, t);
            o[W - 1 - col] = o[W - 1 - col] + t[0];
        }
    }
}

// GB plane rows [lo,hi) = fold(Sobel^T(A,Bv)) ; A,Bv planes start at row ab_lo
template <int VEC>
__device__ __forceinline__ void stage_sobel_adjoint(const EdgeArgs& a, const float* A, const float* Bv, int ab_lo,
                                                    float* GB, int lo, int hi, int G, int tx, int ty) {
    const int W = a.W, H = a.H;
    EE_FOR_TILE(lo, hi) {
        const int col = g * VEC;
        float o[VEC], t[VEC];
        sobel_adj_row<VEC>(A, Bv, ab_lo, row, col, H, W, o);
        if (row == 0) {
            sobel_adj_row<VEC>(A, Bv, ab_lo, -1, col, H, W, t);
#pragma unroll
            for (int k = 0; k < VEC; ++k) o[k] = o[k] + t[k];
        }
        if (row == H - 1) {
            sobel_adj_row<VEC>(A, Bv, ab_lo, H, col, H, W, t);
#pragma unroll
            for (int k = 0; k < VEC; ++k) o[k] = o[k] + t[k];
        }
        st_vec<VEC>(GB + (size_t)(row - lo) * W + col, o);
    }
}

// g_s rows [r0,r1) = fold(Gauss^T(GB)) -> g_x for every channel
template <int VEC, int NC>
__device__ __forceinline__ void stage_gauss_adjoint_store(const EdgeArgs& a, const float* GB, int gb_lo, int b,
                                                          int r0, int r1, int G, int tx, int ty) {
    const int W = a.W, H = a.H;
    const int C = NC ? NC : a.C;
    const size_t hw = (size_t)H * W;
    EE_FOR_TILE(r0, r1) {
        const int col = g * VEC;
        float o[VEC], t[VEC];
        gauss_adj_row<VEC>(a, GB, gb_lo, row, col, H, W, o);
        if (row == 0) {
            gauss_adj_row<VEC>(a, GB, gb_lo, -1, col, H, W, t);
#pragma unroll
            for (int k = 0; k < VEC; ++k) o[k] = o[k] + t[k];
        }
        if (row == H - 1) {
            gauss_adj_row<VEC>(a, GB, gb_lo, H, col, H, W, t);
#pragma unroll
            for (int k = 0; k < VEC; ++k) o[k] = o[k] + t[k];
        }
        for (int c = 0; c < C; ++c) stg_vec<VEC>(a.g_x + at(a.sgx, b, c, row, col), o);
    }
}

// -------------------------------------------------------------------------------------------
// backward.  smem regions: R1 = S then A (TH+8 rows), R2 = Bl then GB (TH+6 rows), R3 = Bv (TH+4)
// -------------------------------------------------------------------------------------------
template <int VEC, int NC, bool BLEND>
__global__ void __launch_bounds__(256) edge_bwd_step125_kernel(const EdgeArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int b = blockIdx.x / a.tiles_per_img;
    const int ti = blockIdx.x - b * a.tiles_per_img;
    const int H = a.H, W = a.W;
    const int C = NC ? NC : a.C;
    const int r0 = ti * a.TH, r1 = min(r0 + a.TH, H);
    const int G = (W + VEC - 1) / VEC;
    const int tx = threadIdx.x % a.GX, ty = threadIdx.x / a.GX;
    const size_t hw = (size_t)H * W;

    float* R1 = smem;
    float* R2 = R1 + (size_t)(a.TH + 8) * W;
    float* R3 = R2 + (size_t)(a.TH + 6) * W;

    const bool want_gx = (a.g_x != nullptr);
    const int s_lo = max(r0 - 4, 0), s_hi = min(r1 + 4, H);
    const int b_lo = max(r0 - 3, 0), b_hi = min(r1 + 3, H);
    // without g_x only the blend adjoint (g_base) is wanted: no halo rows needed
    const int ab_lo = want_gx ? max(r0 - 2, 0) : r0, ab_hi = want_gx ? min(r1 + 2, H) : r1;
    const int gb_lo = max(r0 - 1, 0), gb_hi = min(r1 + 1, H);

    float* S = R1; float* Bl = R2;
    stage_channel_sum<VEC, NC>(a, a.x + (size_t)((int64_t)b * a.sx.b), S, s_lo, s_hi, G, tx, ty);
    __syncthreads();
    stage_blur<VEC>(a, S, s_lo, Bl, b_lo, b_hi, G, tx, ty);
    __syncthreads();

    // A / Bv = dL/dSgx, dL/dSgy on rows [ab_lo, ab_hi)
    float* A = R1; float* Bv = R3;
    EE_FOR_TILE(ab_lo, ab_hi) {
        const int col = g * VEC;
        const size_t pix = (size_t)row * W + col;
        float gx1[VEC], gy1[VEC], mag[VEC], magm[VEC], ge[VEC];
        sobel_at<VEC>(a, Bl, b_lo, row, col, gx1, gy1);
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            mag[k] = magnitude(gx1[k], gy1[k]);
            magm[k] = (mag[k] < a.alpha) ? 0.0f : mag[k];
        }
        if (BLEND) {
            float we[VEC];
#pragma unroll
            for (int k = 0; k < VEC; ++k) we[k] = a.w * to_compare(magm[k], a.high);
            const bool interior = (row >= r0 && row < r1);
            for (int c = 0; c < C; ++c) {
                float bs[VEC], go[VEC], gp[VEC];
                ldg_vec<VEC>(a.base + at(a.sbase, b, c, row, col), bs);
                ldg_vec<VEC>(a.g_in + at(a.sg, b, c, row, col), go);
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    const float pre = bs[k] + we[k];
                    gp[k] = (pre >= 0.0f && pre <= 1.0f) ? go[k] : 0.0f;     // clamp backward, inclusive
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
                float gm = ste_sel(ge[k], magm[k], a.high);              // To_compare.backward
                if (mag[k] < a.alpha) gm = 0.0f;                         // torch.where backward
                mag_backward(gm, mag[k], gx1[k], gy1[k], a.fC, av[k], bv[k], a.nan_compat);
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
