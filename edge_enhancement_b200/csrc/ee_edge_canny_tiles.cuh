// ee_edge_canny_tiles.cuh -- full CannyFilter / CannyFilter_BPDA (+ blend), hysteresis mode, for WIDE images on
// chunk-aligned tiles (same idea as ee_edge_tiles.cuh): 64 x 64 planes = 16 x 16 chunks of 4 x 4 pixels, one per thread
// in every stage, every stage computed on the whole plane with clamped row indices ("garbage" rows / columns outside
// the dependency cone are never read by a valid output).
//   forward : halo 4 rows / 4 columns  (out <- META_nms +-1 <- M +-2 <- Bl +-3 <- S +-4), tile <= 56 x 56
//   backward: halo 8 rows / 8 columns  (out <- GB +-1 <- A +-2 <- META_nms +-3 <- M +-4 <- Bl +-5 <- S +-6), tile <= 48 x 48
// Replaces the guarded strip path of ee_edge_canny_fast.cuh for these shapes; same canonical arithmetic, same bits.
#pragma once
#include "ee_edge_canny_fast.cuh"

namespace ee {

// geometry shared by the two kernels: HALO = 4 (forward) or 8 (backward)
struct TileGeo {
    int H, W, b, r0, r1, c0, c1, a_lo, a_hi, cs, ce, n_ch, Gt, tx, ty, lc, col, ra;
    bool active, p_left, p_right;
};
template <int HALO, int GXT>
__device__ __forceinline__ TileGeo make_tile_geo(const FastArgs& a) {
    TileGeo t;
    t.H = a.e.H; t.W = a.e.W;
    t.b = blockIdx.x / a.e.tiles_per_img;
    const int tq = blockIdx.x - t.b * a.e.tiles_per_img;
    const int ti = tq / a.tiles_x, tj = tq - ti * a.tiles_x;
    t.r0 = ti * a.e.TH; t.r1 = min(t.r0 + a.e.TH, t.H);
    t.c0 = tj * a.TW; t.c1 = min(t.c0 + a.TW, t.W);
    t.a_lo = max(t.r0 - HALO, 0); t.a_hi = min(t.r1 + HALO, t.H);
    t.cs = max(t.c0 - HALO, 0); t.ce = min(t.c1 + HALO, t.W);
    t.n_ch = (t.a_hi - t.a_lo) >> 2; t.Gt = (t.ce - t.cs) >> 2;
    t.tx = threadIdx.x % GXT; t.ty = threadIdx.x / GXT;
    t.active = (t.tx < t.Gt) && (t.ty < t.n_ch);
    t.lc = t.tx * 4; t.col = t.cs + t.lc;
    t.ra = t.a_lo + t.ty * 4;
    t.p_left = (t.lc == 0); t.p_right = (t.lc + 4 == t.ce - t.cs);
    return t;
}

// stages 0-3 (shared): S -> Bl -> (M, META[, gx1, gy1]) -> NMS + thresholds into META.  R1 = S then M, R2 = Bl.
// STORE_G: gx1 goes to GXp; gy1 is returned in registers (the caller stores it into R2 after a barrier).
template <int NC, int R, int Wp, bool STORE_G>
__device__ __forceinline__ void canny_tile_front(const FastArgs& a, const TileGeo& t, float* R1, float* R2, float* META, float* GXp,
                                                 float4 (&gy_keep)[R], const int variant) {
    constexpr int DIVM = (NC == 1) ? 0 : (NC == 3 ? 1 : 2);
    const int C = NC ? NC : a.e.C;
    const size_t hw = (size_t)t.H * t.W;
    auto prow = [&](int r) { return (min(max(r, t.a_lo), t.a_hi - 1) - t.a_lo) * Wp + kPadL + t.lc; };
    // ---- S
    if (t.active) {
        const float* px = a.e.x + (size_t)t.b * C * hw + (size_t)t.ra * t.W + t.col;
        float4 acc[R];
#pragma unroll
        for (int i = 0; i < R; ++i) acc[i] = __ldg(reinterpret_cast<const float4*>(px + i * t.W));
        for (int c = 1; c < C; ++c) {
#pragma unroll
            for (int i = 0; i < R; ++i) acc[i] = f4add(acc[i], __ldg(reinterpret_cast<const float4*>(px + i * t.W + (size_t)c * hw)));
        }
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const float o[4] = {acc[i].x, acc[i].y, acc[i].z, acc[i].w};
            st_plane(R1 + prow(t.ra + i), o, t.p_left, t.p_right, o[0], o[3]);
        }
    }
    __syncthreads();
    // ---- Bl = blur(S)
    if (t.active) {
        const float c0g = a.e.c0, c1g = a.e.c1, c2g = a.e.c2;
        float P[3][4], Q[3][4];
#pragma unroll
        for (int i = 0; i < R + 2; ++i) {
            gauss_partials(ld_win(R1 + prow(t.ra - 1 + i)), c0g, c1g, c2g, P[i % 3], Q[i % 3]);
            if (i >= 2) {
                float o[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) o[k] = (P[(i - 2) % 3][k] + Q[(i - 1) % 3][k]) + P[i % 3][k];
                st_plane(R2 + prow(t.ra + i - 2), o, t.p_left, t.p_right, o[0], o[3]);
            }
        }
    }
    __syncthreads();
    // ---- M (gated magnitude, zero pads), META (direction + 1), optionally gx1 / gy1
    if (t.active) {
        const bool gate = (variant == 1);
        float D[3][4], V[3][4];
#pragma unroll
        for (int i = 0; i < R + 2; ++i) {
            sobel_partials(ld_win(R2 + prow(t.ra - 1 + i)), D[i % 3], V[i % 3]);
            if (i >= 2) {
                float sgx[4], sgy[4], gx1[4], gy1[4], mm[4], mt[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    sgx[k] = fmaf(0.5f, D[(i - 2) % 3][k] + D[i % 3][k], D[(i - 1) % 3][k]);
                    sgy[k] = V[i % 3][k] - V[(i - 2) % 3][k];
                }
                div_channels8<DIVM>(sgx, sgy, a.e.fC, gx1, gy1);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float mag = magnitude(gx1[k], gy1[k]);
                    mm[k] = (gate && mag < a.e.alpha) ? 0.0f : mag;
                    mt[k] = __int_as_float(orient_dir(gx1[k], gy1[k]) + 1);
                }
                const int q = prow(t.ra + i - 2);
                st_plane(R1 + q, mm, t.p_left, t.p_right, 0.0f, 0.0f);
                st_plane(META + q, mt, t.p_left, t.p_right, 0.0f, 0.0f);
                if (STORE_G) {
                    *reinterpret_cast<float4*>(GXp + q) = make_float4(gx1[0], gx1[1], gx1[2], gx1[3]);
                    gy_keep[i - 2] = make_float4(gy1[0], gy1[1], gy1[2], gy1[3]);
                }
            }
        }
    }
    __syncthreads();                        // M complete; every thread has read its Bl windows
    if (STORE_G && t.active) {
#pragma unroll
        for (int i = 0; i < R; ++i) *reinterpret_cast<float4*>(R2 + prow(t.ra + i)) = gy_keep[i];
    }
    // ---- NMS + double threshold into META (rows outside the image contribute zero windows)
    if (t.active) {
        Win wm[3];
#pragma unroll
        for (int i = 0; i < R + 2; ++i) {
            const int rin = t.ra - 1 + i;
            wm[i % 3] = (rin >= 0 && rin < t.H) ? ld_win(R1 + prow(rin)) : zero_win();
            if (i >= 2) {
                float* pmeta = META + prow(t.ra + i - 2);
                const float4 mt = *reinterpret_cast<const float4*>(pmeta);
                float thin[4];
                int meta[4];
                nms_threshold4(a, wm[(i - 2) % 3], wm[(i - 1) % 3], wm[i % 3], mt, thin, meta, variant);
                *reinterpret_cast<float4*>(pmeta) = make_float4(__int_as_float(meta[0]), __int_as_float(meta[1]),
                                                                __int_as_float(meta[2]), __int_as_float(meta[3]));
            }
        }
    }
    __syncthreads();
}

// -------------------------------------------------------------------------------------------
// forward (hysteresis mode): planes R1 = S -> M, R2 = Bl, R3 = META, each (TH + 8) x (64 + 8)
// -------------------------------------------------------------------------------------------
template <int NC, bool BLEND, int R, int PW, int VAR>
__global__ void __launch_bounds__(256, 3) edge_fwd_canny_tiles(const FastArgs a) {
    extern __shared__ __align__(16) float smem[];
    constexpr int Wp = PW + kPadW, GXT = PW / 4;
    const TileGeo t = make_tile_geo<4, GXT>(a);
    const int variant = VAR ? VAR : a.e.variant;
    float* R1 = smem;
    float* R2 = R1 + (a.e.TH + 8) * Wp;
    float* R3 = R2 + (a.e.TH + 8) * Wp;
    float4 unused[R];
    canny_tile_front<NC, R, Wp, false>(a, t, R1, R2, R3, nullptr, unused, variant);
    // ---- hysteresis + emit on the tile's own rows and columns
    const int ro = t.r0 + t.ty * R;
    if (t.tx < t.Gt && ro < t.r1 && t.col >= t.c0 && t.col < t.c1) {
        auto prow = [&](int r) { return (min(max(r, t.a_lo), t.a_hi - 1) - t.a_lo) * Wp + kPadL + t.lc; };
        Geo geo;
        geo.W = t.W; geo.H = t.H;
        int hs[3][4], cw[3][4];
#pragma unroll
        for (int i = 0; i < R + 2; ++i) {
            const int rin = ro - 1 + i;
            meta_partials((rin >= 0 && rin < t.H) ? ld_win(R3 + prow(rin)) : zero_win(), hs[i % 3], cw[i % 3]);
            if (i >= 2) {
                float e[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int n = hs[(i - 2) % 3][k] + hs[(i - 1) % 3][k] + hs[i % 3][k];
                    const int c = cw[(i - 1) % 3][k];
                    const int wih = (meta_lh(c) == 1) && (n >= 2);
                    e[k] = (float)(meta_hi(c) + wih);
                }
                cfast_emit<NC, BLEND, false>(a, geo, t.b, ro + i - 2, t.col, e);
            }
        }
    }
}

// -------------------------------------------------------------------------------------------
// backward (hysteresis mode): R1 = S -> M -> A, R2 = Bl -> gy1 -> GB, R3 = META, R4 = gx1 -> Bv, each (TH + 16) x (64 + 8)
// -------------------------------------------------------------------------------------------
template <int NC, bool BLEND, int R, int PW, int VAR>
__global__ void __launch_bounds__(256, 3) edge_bwd_canny_tiles(const FastArgs a) {
    extern __shared__ __align__(16) float smem[];
    constexpr int Wp = PW + kPadW, GXT = PW / 4;
    const TileGeo t = make_tile_geo<8, GXT>(a);
    const int variant = VAR ? VAR : a.e.variant;
    const int C = NC ? NC : a.e.C;
    const size_t hw = (size_t)t.H * t.W;
    const bool want_gx = (a.e.g_x != nullptr);
    float* R1 = smem;
    float* R2 = R1 + (a.e.TH + 16) * Wp;
    float* R3 = R2 + (a.e.TH + 16) * Wp;
    float* R4 = R3 + (a.e.TH + 16) * Wp;
    auto prow = [&](int r) { return (min(max(r, t.a_lo), t.a_hi - 1) - t.a_lo) * Wp + kPadL + t.lc; };
    float4 gy_keep[R];
    canny_tile_front<NC, R, Wp, true>(a, t, R1, R2, R3, R4, gy_keep, variant);

    // ---- A / Bv on the plane rows: gx1 (R4), gy1 (R2), M (R1) at the own pixels, META windows for the hysteresis vote;
    //      A overwrites M and Bv overwrites gx1 in place.  Only rows [r0-2, r1+2) are needed: the others get zeros and skip
    //      the global loads.
    if (t.active) {
        const float* base_b = a.e.base + (size_t)t.b * C * hw;
        const float* gin_b = a.e.g_in + (size_t)t.b * (BLEND ? C : 1) * hw;
        float* gbase_b = a.e.g_base ? a.e.g_base + (size_t)t.b * C * hw : nullptr;
        const int q_lo = want_gx ? t.r0 - 2 : t.r0, q_hi = want_gx ? t.r1 + 2 : t.r1;
        const bool col_in = (t.col >= t.c0 && t.col < t.c1);
        const float wgt = a.e.w, fC = a.e.fC;
        int hs[3][4], cw[3][4];
#pragma unroll
        for (int i = 0; i < R + 2; ++i) {
            const int rin = t.ra - 1 + i;
            meta_partials((rin >= 0 && rin < t.H) ? ld_win(R3 + prow(rin)) : zero_win(), hs[i % 3], cw[i % 3]);
            if (i >= 2) {
                const int rout = t.ra + i - 2;
                const int q = prow(rout);
                if (!(rout >= q_lo && rout < q_hi)) {
                    if (want_gx) {
                        const float z[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                        st_plane(R1 + q, z, t.p_left, t.p_right, 0.0f, 0.0f);
                        st_plane(R4 + q, z, t.p_left, t.p_right, 0.0f, 0.0f);
                    }
                    continue;
                }
                const int pix = rout * t.W + t.col;
                const float4 tx4 = *reinterpret_cast<const float4*>(R4 + q), ty4 = *reinterpret_cast<const float4*>(R2 + q),
                             tm4 = *reinterpret_cast<const float4*>(R1 + q);
                const float gx1[4] = {tx4.x, tx4.y, tx4.z, tx4.w}, gy1[4] = {ty4.x, ty4.y, ty4.z, ty4.w};
                const float mag[4] = {tm4.x, tm4.y, tm4.z, tm4.w};
                float thin[4], ge[4];
                int meta[4], wih[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    meta[k] = cw[(i - 1) % 3][k];
                    thin[k] = meta_removed(meta[k]) ? 0.0f : mag[k];
                    const int n = hs[(i - 2) % 3][k] + hs[(i - 1) % 3][k] + hs[i % 3][k];
                    wih[k] = (meta_lh(meta[k]) == 1) && (n >= 2);
                }
                if (BLEND) {
                    float we[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) we[k] = wgt * (float)(meta_hi(meta[k]) + wih[k]);
                    const bool interior = col_in && rout >= t.r0 && rout < t.r1;
                    for (int c = 0; c < C; ++c) {
                        const float4 bsc = __ldg(reinterpret_cast<const float4*>(base_b + c * hw + pix));
                        const float4 goc = __ldg(reinterpret_cast<const float4*>(gin_b + c * hw + pix));
                        const float bsv[4] = {bsc.x, bsc.y, bsc.z, bsc.w}, gov[4] = {goc.x, goc.y, goc.z, goc.w};
                        float gp[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float pre = bsv[k] + we[k];
                            gp[k] = (pre >= 0.0f && pre <= 1.0f) ? gov[k] : 0.0f;
                            ge[k] = (c == 0) ? gp[k] * wgt : fmaf(gp[k], wgt, ge[k]);
                        }
                        if (gbase_b && interior)
                            __stcs(reinterpret_cast<float4*>(gbase_b + c * hw + pix), make_float4(gp[0], gp[1], gp[2], gp[3]));
                    }
                } else {
                    const float4 tg = __ldg(reinterpret_cast<const float4*>(gin_b + pix));
                    ge[0] = tg.x; ge[1] = tg.y; ge[2] = tg.z; ge[3] = tg.w;
                }
                if (want_gx) {
                    float av[4], bv[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        float gm = g_thin_of_v(variant, a.e.low, a.e.high, MODE_HYST, ge[k], thin[k], wih[k]);
                        if (meta_removed(meta[k])) gm = 0.0f;
                        mag_backward(gm, mag[k], gx1[k], gy1[k], fC, av[k], bv[k]);
                    }
                    st_plane(R1 + q, av, t.p_left, t.p_right, 0.0f, 0.0f);
                    st_plane(R4 + q, bv, t.p_left, t.p_right, 0.0f, 0.0f);
                }
            }
        }
    }
    if (!want_gx) return;
    __syncthreads();

    // ---- GB = fold(Sobel^T(A, Bv)) on the plane rows -> R2
    const AdjBorder bd = {t.col == 0, t.col + 4 == t.W};
    const bool ring = bd.left || bd.right;
    if (t.active) {
        float HA[3][4], HB[3][4], HAr[3], HBr[3];
        auto loadp = [&](int i, int rin, bool valid) {
            if (valid) {
                const int q = prow(rin);
                sobel_adj_partials(ld_win(R1 + q), ld_win(R4 + q), bd, HA[i % 3], HB[i % 3], HAr[i % 3], HBr[i % 3]);
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
                const float tt = xa + yb;
                if (bd.left) o[0] = o[0] + tt; else o[3] = o[3] + tt;
            }
        };
        auto store = [&](int row, const float (&o)[4]) { st_plane(R2 + prow(row), o, t.p_left, t.p_right, 0.0f, 0.0f); };
        adj_chunk_even<R>(t.ra, t.H, loadp, combine, store);
    }
    __syncthreads();

    // ---- g_s = fold(Gauss^T(GB)) on the tile's own rows and columns -> every channel of g_x
    const int ro = t.r0 + t.ty * R;
    if (t.tx < t.Gt && ro < t.r1 && t.col >= t.c0 && t.col < t.c1) {
        const float c0g = a.e.c0, c1g = a.e.c1, c2g = a.e.c2;
        float* gx_b = a.e.g_x + (size_t)t.b * C * hw;
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
                const float tt = (Pr[(i - 2) % 3] + Qr[(i - 1) % 3]) + Pr[i % 3];
                if (bd.left) o[0] = o[0] + tt; else o[3] = o[3] + tt;
            }
        };
        auto store = [&](int row, const float (&o)[4]) {
            const float4 v = make_float4(o[0], o[1], o[2], o[3]);
            float* pg = gx_b + (size_t)row * t.W + t.col;
            for (int c = 0; c < C; ++c) __stcs(reinterpret_cast<float4*>(pg + c * hw), v);
        };
        adj_chunk_even<R>(ro, t.H, loadp, combine, store);
    }
}

}  // namespace ee
