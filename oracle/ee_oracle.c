/*
 * ee_oracle.c -- CPU ORACLE for the edge-enhancement + PGD-step hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this file's shared object.  The product (edge_enhancement_b200/) never does and has no
 * CPU fallback.
 *
 * What it is: a plain-C, whole-image, full-plane restatement of the reference's algorithm
 * (Aiqz/Edge-Enhancement, utils/core.py and utils/attacks.py).  Every function cites the
 * reference file:line it follows.  It is written independently of the CUDA kernels
 * (image-level loops over complete planes, no tiling) but evaluates the SAME canonical
 * fp32 expression trees (DESIGN.md "Canonical arithmetic"), so CUDA-vs-oracle comparisons
 * are bit-exact, while oracle-vs-reference comparisons (tests/golden, made by running the
 * real reference modules) are within the north-star tolerances.
 *
 * Parity pin: the reference has no tests/golden vectors of its own, so
 * the oracle is pinned against outputs of the reference itself run in the build container
 * (oracle/make_golden.py -> tests/golden/ *.npz).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math [-mfma] -fopenmp -shared -fPIC
 *        (-ffp-contract=off is REQUIRED: fused multiply-adds happen only where fmaf() is
 *        written.)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EE_STEP125 0
#define EE_CANNY   1
#define EE_BPDA    2

typedef struct {
    int   variant;            /* EE_STEP125 / EE_CANNY / EE_BPDA                           */
    float c0, c1, c2;         /* 3x3 Gaussian: corner, edge, centre (core.py:58-72, :164)  */
    float alpha;              /* magnitude gate (core.py:263-264, :574-575); unused by BPDA */
    float low_thr, high_thr;  /* thresholds, already rounded to fp32 like torch does       */
    int   has_low, has_high;  /* "is not None" flags of forward() (core.py:295,309)        */
    int   hysteresis;         /* core.py:317                                               */
    int   nan_compat;         /* backward: NaN where the magnitude is 0, like autograd through (gx^2+gy^2)**0.5 (core.py:250, :453, :571) */
} ee_oracle_params;

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* To_compare.forward, core.py:338-347:  out = in.clone(); out[out <= thr] = 0; out[out > thr] = 1.
 * The two masked writes run in that order, so with a NEGATIVE threshold the zeros written by
 * the first are turned into ones by the second (0 > thr).  NaN stays NaN. */
static inline float to_compare(float v, float thr)
{
    if (v > thr) return 1.0f;
    if (v <= thr) return (0.0f > thr) ? 1.0f : 0.0f;
    return v;
}
/* (sign(v - thr) + 1) / 2 with safeSign(0) = -1, core.py:115-118, :299-310. */
static inline float sign_step(float v, float thr) { return (v - thr > 0.0f) ? 1.0f : 0.0f; }

/* ------------------------------------------------------------------------------------
 * Stage A: channel sum.  The reference blurs every channel (core.py:560-563) and lets the
 * Sobel conv sum over channels (weight.repeat(1,C,1,1), core.py:566-567).  Blur, replicate
 * padding and the Sobel conv are all linear, so summing the channels FIRST is the same
 * function; it is the canonical order here: s = ((x0 + x1) + x2) + ...
 * ---------------------------------------------------------------------------------- */
static void channel_sum(const float *x, int C, int H, int W, float *s)
{
    const size_t hw = (size_t)H * W;
    for (size_t i = 0; i < hw; ++i) {
        float acc = x[i];
        for (int c = 1; c < C; ++c) acc = acc + x[(size_t)c * hw + i];
        s[i] = acc;
    }
}

/* ------------------------------------------------------------------------------------
 * Stage B: 3x3 Gaussian with ReplicationPad2d(1)  (core.py:162, :233-236, :560-563).
 * G = [[c0,c1,c0],[c1,c2,c1],[c0,c1,c0]] (symmetric because it depends on the distance to
 * the centre only, core.py:62-66).  Canonical order:
 *     e = left + right;  P = fma(c1, mid, c0*e);  Q = fma(c2, mid, c1*e)
 *     blur(i,j) = (P(i-1,j) + Q(i,j)) + P(i+1,j)          rows/cols clamped (replicate)
 * ---------------------------------------------------------------------------------- */
static void blur3_replicate(const float *s, int H, int W, float c0, float c1, float c2,
                            float *P, float *Q, float *out)
{
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            float l = s[(size_t)i * W + clampi(j - 1, 0, W - 1)];
            float r = s[(size_t)i * W + clampi(j + 1, 0, W - 1)];
            float m = s[(size_t)i * W + j];
            float e = l + r;
            P[(size_t)i * W + j] = fmaf(c1, m, c0 * e);
            Q[(size_t)i * W + j] = fmaf(c2, m, c1 * e);
        }
    for (int i = 0; i < H; ++i) {
        int iu = clampi(i - 1, 0, H - 1), id = clampi(i + 1, 0, H - 1);
        for (int j = 0; j < W; ++j)
            out[(size_t)i * W + j] = (P[(size_t)iu * W + j] + Q[(size_t)i * W + j]) + P[(size_t)id * W + j];
    }
}

/* ------------------------------------------------------------------------------------
 * Stage C: Sobel x / y with ReplicationPad2d(1) on the BLURRED image (core.py:176, :250-252,
 * :565-567).  Kx = [[-.5,0,.5],[-1,0,1],[-.5,0,.5]] (core.py:75-84), Ky = Kx^T (core.py:180).
 *     D(i,j) = b(i,j+1) - b(i,j-1)                 V(i,j) = fma(.5, b(i,j-1)+b(i,j+1), b(i,j))
 *     Sgx    = fma(.5, D(i-1,j)+D(i+1,j), D(i,j))  Sgy    = V(i+1,j) - V(i-1,j)
 * then /C (true division, core.py:256,446,570).
 * ---------------------------------------------------------------------------------- */
static void sobel3_replicate(const float *b, int H, int W, int C, float *D, float *V,
                             float *gx1, float *gy1)
{
    const float fC = (float)C;
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            float l = b[(size_t)i * W + clampi(j - 1, 0, W - 1)];
            float r = b[(size_t)i * W + clampi(j + 1, 0, W - 1)];
            float m = b[(size_t)i * W + j];
            D[(size_t)i * W + j] = r - l;
            V[(size_t)i * W + j] = fmaf(0.5f, l + r, m);
        }
    for (int i = 0; i < H; ++i) {
        int iu = clampi(i - 1, 0, H - 1), id = clampi(i + 1, 0, H - 1);
        for (int j = 0; j < W; ++j) {
            float sgx = fmaf(0.5f, D[(size_t)iu * W + j] + D[(size_t)id * W + j], D[(size_t)i * W + j]);
            float sgy = V[(size_t)id * W + j] - V[(size_t)iu * W + j];
            gx1[(size_t)i * W + j] = sgx / fC;
            gy1[(size_t)i * W + j] = sgy / fC;
        }
    }
}

/* Orientation of core.py:258-260,270:  atan(gy/gx)*(360/pi)+180 -> round(/45)*45 -> (/45)%8, of which
 * NMS only uses the direction pair `bin mod 4` (core.py:275-281).  With v = atan(r)*8/pi + 4, r = gy/gx,
 * bin = round(v) mod 8 and the bin boundaries are at |r| = t_i = tan((2i+1)*pi/16), i = 0..3:
 *     m = #{ i : |r| > t_i } ;  dir = (r > 0) ? m mod 4 : (4 - m) mod 4
 * Canonical form (no atan, no division): with a = |gx|, s = sign(gx)*gy,  |r| > t_i  <=>  |s| > t_i*a.
 * gx = 0, gy != 0 gives r = +-inf -> dir 0, as in the reference (atan(+-inf) -> bin 0 or 8).  gx = gy = 0
 * (the reference's NaN orientation, no direction) has magnitude 0, so its direction is immaterial. */
static const float EE_TAN_T[4] = { 0.19891236737965800691f, 0.66817863791929891999f,
                                   1.49660576266548901760f, 5.02733949212584810451f };

static inline int orient_dir(float gx1, float gy1)
{
    const float a = fabsf(gx1);
    const float s = (gx1 < 0.0f) ? -gy1 : gy1;
    const float as = fabsf(s);
    int m = 0;
    for (int i = 0; i < 4; ++i) m += (as > EE_TAN_T[i] * a);
    return (s > 0.0f) ? (m & 3) : ((4 - m) & 3);
}

/* (row, col) offset of the -1 tap of directional kernel k (core.py:87-112; cv2 output). */
static const int EE_DIR_DR[8] = { 0, -1, -1, -1, 0, 1, 1, 1 };
static const int EE_DIR_DC[8] = { 1, 1, 0, -1, -1, -1, 0, 1 };

typedef struct {
    float *s, *P, *Q, *blur, *D, *V, *gx1, *gy1, *mag, *magm, *thin, *t, *hi;
    unsigned char *removed, *wih;
} planes_t;

static int planes_alloc(planes_t *p, size_t hw)
{
    float *buf = (float *)malloc(sizeof(float) * hw * 13);
    if (!buf) return -1;
    p->s = buf; p->P = buf + hw; p->Q = buf + 2 * hw; p->blur = buf + 3 * hw; p->D = buf + 4 * hw;
    p->V = buf + 5 * hw; p->gx1 = buf + 6 * hw; p->gy1 = buf + 7 * hw; p->mag = buf + 8 * hw;
    p->magm = buf + 9 * hw; p->thin = buf + 10 * hw; p->t = buf + 11 * hw; p->hi = buf + 12 * hw;
    p->removed = (unsigned char *)malloc(hw * 2);
    if (!p->removed) { free(buf); return -1; }
    p->wih = p->removed + hw;
    return 0;
}
static void planes_free(planes_t *p) { free(p->s); free(p->removed); }

/* Forward of one image: fills every intermediate plane and the edge map.
 * STEP125: core.py:549-585.  CANNY: core.py:222-326.  BPDA: core.py:426-505. */
static void edge_forward_image(const float *x, int C, int H, int W, const ee_oracle_params *pr,
                               planes_t *pl, float *edge)
{
    const size_t hw = (size_t)H * W;
    channel_sum(x, C, H, W, pl->s);
    blur3_replicate(pl->s, H, W, pr->c0, pr->c1, pr->c2, pl->P, pl->Q, pl->blur);
    sobel3_replicate(pl->blur, H, W, C, pl->D, pl->V, pl->gx1, pl->gy1);
    for (size_t i = 0; i < hw; ++i) {
        float gx = pl->gx1[i], gy = pl->gy1[i];
        float m = sqrtf(gx * gx + gy * gy);               /* core.py:257,447,571 */
        pl->mag[i] = m;
        if (pr->variant == EE_BPDA) pl->magm[i] = m;      /* BPDA has no alpha gate */
        else pl->magm[i] = (m < pr->alpha) ? 0.0f : m;    /* core.py:263-264,574-575 */
    }
    if (pr->variant == EE_STEP125) {
        /* To_compare.forward, core.py:338-347 (strict >) ; core.py:578-583 */
        for (size_t i = 0; i < hw; ++i) edge[i] = to_compare(pl->magm[i], pr->high_thr);
        return;
    }
    /* non-maximum suppression, core.py:268-290 / :455-480 (directional conv is zero padded) */
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            size_t q = (size_t)i * W + j;
            int rem = 0;
            {
                int k = orient_dir(pl->gx1[q], pl->gy1[q]);
                float m = pl->magm[q];
                int i1 = i + EE_DIR_DR[k], j1 = j + EE_DIR_DC[k];
                int i2 = i + EE_DIR_DR[k + 4], j2 = j + EE_DIR_DC[k + 4];
                float n1 = (i1 >= 0 && i1 < H && j1 >= 0 && j1 < W) ? pl->magm[(size_t)i1 * W + j1] : 0.0f;
                float n2 = (i2 >= 0 && i2 < H && j2 >= 0 && j2 < W) ? pl->magm[(size_t)i2 * W + j2] : 0.0f;
                float d1 = m - n1, d2 = m - n2;
                float mn = d1 < d2 ? d1 : d2;
                rem = !(mn > 0.0f);
            }
            pl->removed[q] = (unsigned char)rem;
            pl->thin[q] = rem ? 0.0f : pl->magm[q];
        }
    const int bpda = (pr->variant == EE_BPDA);
    /* core.py:326: no low threshold -> the thinned magnitude itself.  CannyFilter_BPDA has no
     * `else: thin_edges = low` branch (core.py:482-505), so it ALSO returns the raw thinned
     * magnitude when only the low threshold is given. */
    if (!pr->has_low || (bpda && !pr->has_high)) { memcpy(edge, pl->thin, hw * sizeof(float)); return; }
    if (!pr->has_high) {                                                        /* core.py:323-324 */
        for (size_t i = 0; i < hw; ++i) edge[i] = sign_step(pl->thin[i], pr->low_thr);
        return;
    }
    for (size_t i = 0; i < hw; ++i) {   /* core.py:300,310,315 / :486-492 */
        float lo = bpda ? to_compare(pl->thin[i], pr->low_thr) : sign_step(pl->thin[i], pr->low_thr);
        float hi = bpda ? to_compare(pl->thin[i], pr->high_thr) : sign_step(pl->thin[i], pr->high_thr);
        pl->hi[i] = hi;
        pl->t[i] = lo * 0.5f + hi * 0.5f;
    }
    if (!pr->hysteresis) { memcpy(edge, pl->t, hw * sizeof(float)); return; }
    /* hysteresis, core.py:317-321 / :494-503: weak = (t == .5); keep if 1.25*sum3x3(t) > 1 */
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            size_t q = (size_t)i * W + j;
            float acc = 0.0f;
            for (int a = -1; a <= 1; ++a)
                for (int b = -1; b <= 1; ++b) {
                    int ii = i + a, jj = j + b;
                    if (ii >= 0 && ii < H && jj >= 0 && jj < W) acc += 1.25f * pl->t[(size_t)ii * W + jj];
                }
            int weak = (pl->t[q] == 0.5f);
            int wih = weak && (acc > 1.0f);
            pl->wih[q] = (unsigned char)wih;
            edge[q] = pl->hi[q] + (wih ? 1.0f : 0.0f);
        }
}

/* edge [B,1,H,W] = filter(x [B,C,H,W]) */
int ee_oracle_edge_fwd(const float *x, float *edge, int B, int C, int H, int W,
                       const ee_oracle_params *pr)
{
    const size_t hw = (size_t)H * W;
    int err = 0;
#pragma omp parallel
    {
        planes_t pl;
        int ok = planes_alloc(&pl, hw) == 0;
        if (!ok) {
#pragma omp atomic write
            err = -1;
        }
#pragma omp for schedule(static)
        for (int b = 0; b < B; ++b)
            if (ok) edge_forward_image(x + (size_t)b * C * hw, C, H, W, pr, &pl, edge + (size_t)b * hw);
        if (ok) planes_free(&pl);
    }
    return err;
}

/* Blend, e.g. Tiny_ImageNet/models_tinyimagenet/resnet_EE.py:189-191:
 *     out_c = clamp(base_c + w*edge, 0, 1)      (NaN propagates through torch.clamp) */
static inline float clamp01_nan(float v) { return (v != v) ? v : (v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v)); }

int ee_oracle_edge_blend_fwd(const float *x, const float *base, float *out, float *edge_or_null,
                             int B, int C, int H, int W, const ee_oracle_params *pr, float w)
{
    const size_t hw = (size_t)H * W;
    int err = 0;
#pragma omp parallel
    {
        planes_t pl;
        float *edge = (float *)malloc(hw * sizeof(float));
        int ok = (planes_alloc(&pl, hw) == 0) && edge;
        if (!ok) {
#pragma omp atomic write
            err = -1;
        }
#pragma omp for schedule(static)
        for (int b = 0; b < B; ++b) {
            if (!ok) continue;
            edge_forward_image(x + (size_t)b * C * hw, C, H, W, pr, &pl, edge);
            if (edge_or_null) memcpy(edge_or_null + (size_t)b * hw, edge, hw * sizeof(float));
            for (int c = 0; c < C; ++c) {
                const float *bs = base + ((size_t)b * C + c) * hw;
                float *o = out + ((size_t)b * C + c) * hw;
                for (size_t i = 0; i < hw; ++i) {
                    float we = w * edge[i];
                    o[i] = clamp01_nan(bs[i] + we);
                }
            }
        }
        free(edge);
        if (ok) planes_free(&pl);
    }
    return err;
}

/* ------------------------------------------------------------------------------------
 * Backward.  g_thin from the gradient w.r.t. the module output, per variant / mode.
 *   STEP125: To_compare.backward, core.py:350-358:      g * [thr < in <= 1.001]
 *   CANNY  : BinaryConnectDeterministic.backward core.py:138-145 through (sign(.)+1)/2:
 *            0.5 * g * [|thin - thr| <= 1.001]; with hysteresis only `high` carries gradient
 *            (core.py:319-321: weak / weak_is_high are integer tensors).
 *   BPDA   : To_compare / To_eq (core.py:482-503); closed form in SURVEY.md A.3.
 * ---------------------------------------------------------------------------------- */
/* masked assignments of the reference (grad_input[mask] = 0): a SELECT, so a non-finite upstream
 * gradient outside the window becomes 0, not NaN */
static inline float ste_sel(float g, float v, float thr) { return (v > thr && v <= 1.001f) ? g : 0.0f; }
static inline float bcd_sel(float g, float v, float thr) { return (fabsf(v - thr) > 1.001f) ? 0.0f : g; }

static float g_thin_of(const ee_oracle_params *pr, const planes_t *pl, size_t q, float ge)
{
    float th = pl->thin[q];
    if (!pr->has_low) return ge;
    if (pr->variant == EE_CANNY) {
        if (!pr->has_high) return bcd_sel(0.5f * ge, th, pr->low_thr);
        if (!pr->hysteresis) {
            float h = 0.5f * ge;   /* d(low*.5 + high*.5) */
            return bcd_sel(0.5f * h, th, pr->low_thr) + bcd_sel(0.5f * h, th, pr->high_thr);
        }
        return bcd_sel(0.5f * ge, th, pr->high_thr);
    }
    /* BPDA */
    if (!pr->has_high) return ge;           /* raw thinned magnitude is returned, see forward */
    if (!pr->hysteresis) {
        float h = 0.5f * ge;
        return ste_sel(h, th, pr->low_thr) + ste_sel(h, th, pr->high_thr);
    }
    {
        float wih = pl->wih[q] ? 1.0f : 0.0f;
        float gt = ge * wih;                        /* through To_eq(t) * weak_1 */
        float g_low = 0.5f * gt;
        float g_high = ge + 0.5f * gt;
        return ste_sel(g_low, th, pr->low_thr) + ste_sel(g_high, th, pr->high_thr);
    }
}

/* Adjoint of Stage C + Stage B for one image.  a = dL/dSgx, b = dL/dSgy on the image domain.
 * corr^T into the padded frame followed by the replicate-pad fold, written as
 * a gather with a fixed evaluation order:
 *   HA(r,q) = a(r,q-1) - a(r,q+1)                       (zero extended)
 *   HB(r,q) = fma(.5, b(r,q-1)+b(r,q+1), b(r,q))        (zero extended)
 *   T(p,q)  = fma(.5, HA(p-1,q)+HA(p+1,q), HA(p,q)) + (HB(p-1,q) - HB(p+1,q)),  p in [-1,H], q in [-1,W]
 *   column fold: F(p,j) = T(p,j) [+ T(p,-1) if j==0] [+ T(p,W) if j==W-1]
 *   row fold   : gb(i,j) = F(i,j) [+ F(-1,j) if i==0] [+ F(H,j) if i==H-1]
 * and the same with the Gaussian for the blur stage. */
static inline float at0(const float *p, int H, int W, int i, int j)
{
    return (i >= 0 && i < H && j >= 0 && j < W) ? p[(size_t)i * W + j] : 0.0f;
}

static float sobelT_T(const float *a, const float *b, int H, int W, int p, int q)
{
    float ha[3], hb[3];
    for (int d = -1; d <= 1; ++d) {
        int r = p + d;
        ha[d + 1] = at0(a, H, W, r, q - 1) - at0(a, H, W, r, q + 1);
        hb[d + 1] = fmaf(0.5f, at0(b, H, W, r, q - 1) + at0(b, H, W, r, q + 1), at0(b, H, W, r, q));
    }
    float xa = fmaf(0.5f, ha[0] + ha[2], ha[1]);
    float yb = hb[0] - hb[2];
    return xa + yb;
}

static float gaussT_T(const float *g, int H, int W, float c0, float c1, float c2, int p, int q)
{
    float P[3], Qm = 0.0f;
    for (int d = -1; d <= 1; ++d) {
        int r = p + d;
        float l = at0(g, H, W, r, q - 1), rr = at0(g, H, W, r, q + 1), m = at0(g, H, W, r, q);
        float e = l + rr;
        P[d + 1] = fmaf(c1, m, c0 * e);
        if (d == 0) Qm = fmaf(c2, m, c1 * e);
    }
    return (P[0] + Qm) + P[2];
}

typedef float (*tapfn)(const void *ctx, int p, int q);

static float fold_cols(tapfn f, const void *ctx, int W, int p, int j)
{
    float v = f(ctx, p, j);
    if (j == 0) v = v + f(ctx, p, -1);
    if (j == W - 1) v = v + f(ctx, p, W);
    return v;
}
static float fold_all(tapfn f, const void *ctx, int H, int W, int i, int j)
{
    float v = fold_cols(f, ctx, W, i, j);
    if (i == 0) v = v + fold_cols(f, ctx, W, -1, j);
    if (i == H - 1) v = v + fold_cols(f, ctx, W, H, j);
    return v;
}

typedef struct { const float *a, *b; int H, W; } sob_ctx;
typedef struct { const float *g; int H, W; float c0, c1, c2; } gau_ctx;
static float sob_tap(const void *c, int p, int q) { const sob_ctx *s = (const sob_ctx *)c; return sobelT_T(s->a, s->b, s->H, s->W, p, q); }
static float gau_tap(const void *c, int p, int q) { const gau_ctx *s = (const gau_ctx *)c; return gaussT_T(s->g, s->H, s->W, s->c0, s->c1, s->c2, p, q); }

/* ge[H*W] = dL/d(edge) ; writes gs[H*W] = dL/d(s) (identical for every channel of x). */
static void edge_backward_image(const float *ge, int C, int H, int W, const ee_oracle_params *pr,
                                const planes_t *pl, float *a, float *b, float *gb, float *gs)
{
    const size_t hw = (size_t)H * W;
    const float fC = (float)C;
    for (size_t q = 0; q < hw; ++q) {
        float gm;
        if (pr->variant == EE_STEP125) {
            gm = ste_sel(ge[q], pl->magm[q], pr->high_thr);
            if (pl->mag[q] < pr->alpha) gm = 0.0f;              /* torch.where backward */
        } else {
            gm = g_thin_of(pr, pl, q, ge[q]);
            if (pl->removed[q]) gm = 0.0f;                       /* core.py:290 / :480 */
            if (pr->variant == EE_CANNY && pl->mag[q] < pr->alpha) gm = 0.0f;
        }
        /* mag = u^0.5, u = gx1^2 + gy1^2 ; autograd evaluates g*0.5*u^-0.5, then *2*gx1, then /C
         * (three divisions).  Canonical form here: one division, dL/dSgx = (g / (mag*C)) * gx1 --
         * the same value up to ~2 ulp, well inside the 1e-5 gradient tolerance.
         * Sub-gradient at mag == 0 defined as 0 (the reference yields NaN). */
        float m = pl->mag[q];
        /* nan_compat: what autograd really computes at u = 0: g * 0.5 * u^-0.5 = g * inf (NaN for g == 0), times
         * 2*gx1 = 0 -> NaN whatever g is; the two adjoint stencils below then spread it over the 5 x 5 neighbourhood */
        if (pr->nan_compat && m == 0.0f) { a[q] = NAN; b[q] = NAN; continue; }
        if (gm == 0.0f || m == 0.0f) { a[q] = 0.0f; b[q] = 0.0f; continue; }
        float t = gm / (m * fC);
        a[q] = t * pl->gx1[q];
        b[q] = t * pl->gy1[q];
    }
    sob_ctx sc = { a, b, H, W };
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) gb[(size_t)i * W + j] = fold_all(sob_tap, &sc, H, W, i, j);
    gau_ctx gc = { gb, H, W, pr->c0, pr->c1, pr->c2 };
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) gs[(size_t)i * W + j] = fold_all(gau_tap, &gc, H, W, i, j);
}

/* g_x [B,C,H,W] from g_edge [B,1,H,W] (module-level backward of CannyFilter*.forward). */
int ee_oracle_edge_bwd(const float *g_edge, const float *x, float *g_x, int B, int C, int H, int W,
                       const ee_oracle_params *pr)
{
    const size_t hw = (size_t)H * W;
    int err = 0;
#pragma omp parallel
    {
        planes_t pl;
        float *tmp = (float *)malloc(hw * sizeof(float) * 5);
        int ok = (planes_alloc(&pl, hw) == 0) && tmp;
        if (!ok) {
#pragma omp atomic write
            err = -1;
        }
#pragma omp for schedule(static)
        for (int bi = 0; bi < B; ++bi) {
            if (!ok) continue;
            float *edge = tmp, *a = tmp + hw, *b = tmp + 2 * hw, *gb = tmp + 3 * hw, *gs = tmp + 4 * hw;
            edge_forward_image(x + (size_t)bi * C * hw, C, H, W, pr, &pl, edge);
            edge_backward_image(g_edge + (size_t)bi * hw, C, H, W, pr, &pl, a, b, gb, gs);
            for (int c = 0; c < C; ++c) memcpy(g_x + ((size_t)bi * C + c) * hw, gs, hw * sizeof(float));
        }
        free(tmp);
        if (ok) planes_free(&pl);
    }
    return err;
}

/* Fused backward of out = clamp(base + w*edge(x), 0, 1):
 *   g_pre_c = g_out_c * [0 <= pre_c <= 1]  (clamp backward, inclusive) ; g_base_c = g_pre_c
 *   g_e     = sum_c (g_pre_c * w)          (mul backward then broadcast-sum over channels)
 *   g_x     = edge_backward(g_e)           (edge path only; HFS^T(g_base) is added by autograd) */
int ee_oracle_edge_blend_bwd(const float *g_out, const float *x, const float *base, float *g_x,
                             float *g_base, int B, int C, int H, int W, const ee_oracle_params *pr, float w)
{
    const size_t hw = (size_t)H * W;
    int err = 0;
#pragma omp parallel
    {
        planes_t pl;
        float *tmp = (float *)malloc(hw * sizeof(float) * 6);
        int ok = (planes_alloc(&pl, hw) == 0) && tmp;
        if (!ok) {
#pragma omp atomic write
            err = -1;
        }
#pragma omp for schedule(static)
        for (int bi = 0; bi < B; ++bi) {
            if (!ok) continue;
            float *edge = tmp, *a = tmp + hw, *b = tmp + 2 * hw, *gb = tmp + 3 * hw, *gs = tmp + 4 * hw, *ge = tmp + 5 * hw;
            edge_forward_image(x + (size_t)bi * C * hw, C, H, W, pr, &pl, edge);
            for (size_t q = 0; q < hw; ++q) {
                float acc = 0.0f;
                for (int c = 0; c < C; ++c) {
                    size_t o = ((size_t)bi * C + c) * hw + q;
                    float we = w * edge[q];
                    float pre = base[o] + we;
                    float gp = (pre >= 0.0f && pre <= 1.0f) ? g_out[o] : 0.0f;
                    if (g_base) g_base[o] = gp;
                    acc = (c == 0) ? gp * w : fmaf(gp, w, acc);
                }
                ge[q] = acc;
            }
            if (g_x) {
                edge_backward_image(ge, C, H, W, pr, &pl, a, b, gb, gs);
                for (int c = 0; c < C; ++c) memcpy(g_x + ((size_t)bi * C + c) * hw, gs, hw * sizeof(float));
            }
        }
        free(tmp);
        if (ok) planes_free(&pl);
    }
    return err;
}

/* ------------------------------------------------------------------------------------
 * Attack updates (utils/attacks.py).  torch.sign: sign(0) = 0, sign(NaN) = 0 (probed on
 * torch 2.11: (0<g)-(g<0)).  torch.min/max/clamp propagate NaN.
 * ---------------------------------------------------------------------------------- */
static inline float sgn(float g) { return (float)((g > 0.0f) - (g < 0.0f)); }
static inline float maxn(float a, float b) { return (a != a) ? a : ((b != b) ? b : (a > b ? a : b)); }
static inline float minn(float a, float b) { return (a != a) ? a : ((b != b) ? b : (a < b ? a : b)); }

/* PGD L-inf step, attacks.py:25-27 (also :52-54, :82-84, :257-259, :298-300, :318-320,
 * :353-355, :414-416, :466-468, :505-507 with alpha_signed = -step for the targeted ones). */
void ee_oracle_pgd_linf_step(const float *x, const float *g, const float *x0, float *out, int64_t n,
                             float alpha_signed, float eps, float lo, float hi)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        float t = x[i] + alpha_signed * sgn(g[i]);
        t = minn(maxn(t, x0[i] - eps), x0[i] + eps);
        out[i] = minn(maxn(t, lo), hi);
    }
}

/* FGSM, attacks.py:121-126: one signed step and the [0,1] clamp, no eps projection. */
void ee_oracle_fgsm_step(const float *x, const float *g, float *out, int64_t n, float alpha_signed,
                         float lo, float hi)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        float t = x[i] + alpha_signed * sgn(g[i]);
        out[i] = minn(maxn(t, lo), hi);
    }
}

/* free / fast AT delta update, ImageNet/free_imagenet/AT_hfs_canny_free_imagenet_ddp.py:330-332
 * then :314-315 of the next repeat:  delta += a*sign(g); delta.clamp_(-eps,eps);
 * x_adv = clamp(x0 + delta, 0, 1). */
void ee_oracle_free_at_step(float *delta, const float *g, const float *x0, float *x_adv, int64_t n,
                            float alpha, float eps, float lo, float hi)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        float d = delta[i] + alpha * sgn(g[i]);
        d = minn(maxn(d, -eps), eps);
        delta[i] = d;
        if (x_adv) x_adv[i] = minn(maxn(x0[i] + d, lo), hi);
    }
}

/* CW L-inf inner update, attacks.py:213-222. */
void ee_oracle_cw_linf_step(const float *adv, const float *g, const float *x, const float *min_x,
                            const float *max_x, float *out, int64_t n, float step, float magnitude)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        float t = adv[i] + step * sgn(g[i]);
        t = maxn(minn(t, x[i] + magnitude), x[i] - magnitude);
        t = minn(maxn(t, 0.0f), 1.0f);
        out[i] = maxn(minn(t, max_x[i]), min_x[i]);
    }
}

/* Per-sample RMS "l2_norm" of attacks.py:360-366: sqrt(mean(v^2)).  Canonical reduction
 * order (shared with the CUDA kernel): 1024 lane partials, lane l sums elements l, l+1024,
 * ... in index order; then a fixed pairwise tree: strides 16..1 inside every group of 32
 * lanes, then strides 16..1 over the 32 group sums. */
#define EE_L2_LANES 1024
static float rms_canonical(const float *v, int64_t n, const float *v2_sub, int mode)
{
    /* mode 0: v ; mode 1: v - v2_sub */
    float part[EE_L2_LANES];
    for (int l = 0; l < EE_L2_LANES; ++l) {
        float acc = 0.0f;
        for (int64_t i = l; i < n; i += EE_L2_LANES) {
            float e = mode ? (v[i] - v2_sub[i]) : v[i];
            acc = fmaf(e, e, acc);
        }
        part[l] = acc;
    }
    /* warp-first tree: inside each group of 32 lanes strides 16..1, then the 32 group sums */
    float wsum[32];
    for (int w = 0; w < EE_L2_LANES / 32; ++w) {
        float *p = part + 32 * w;
        for (int s = 16; s >= 1; s >>= 1)
            for (int l = 0; l < s; ++l) p[l] = p[l] + p[l + s];
        wsum[w] = p[0];
    }
    for (int s = 16; s >= 1; s >>= 1)
        for (int l = 0; l < s; ++l) wsum[l] = wsum[l] + wsum[l + s];
    return sqrtf(wsum[0] / (float)n);
}

/* The one-pass CUDA kernel (edge_enhancement_b200/csrc/ee_pgd_l2.cuh) keeps a sample on chip in a cluster of K CTAs of T
 * threads; its reduction order, restated: the sample's float4 words are cut into K slices of slice4 words; inside a slice
 * thread t accumulates words t, t + T, ... element by element with fmaf; shuffle tree (strides 16..1) inside each warp;
 * the T/32 warp sums padded with zeros to 32 and the same tree; then the K slice sums left to right. */
#define EE_L2C_MAX_THREADS 512
#define EE_L2C_ONE_CTA4 3072
#define EE_L2C_TARGET_SLICE4 4096
#define EE_L2C_MAX_CLUSTER 8
static int l2_cluster_plan(int64_t n_per, int *K, int64_t *slice4, int *threads)
{
    if (n_per <= 0 || (n_per & 3)) return 0;
    int64_t n4 = n_per >> 2;
    int k = 1;
    if (n4 > EE_L2C_ONE_CTA4)
        while (k < EE_L2C_MAX_CLUSTER && (n4 + k - 1) / k > EE_L2C_TARGET_SLICE4) k <<= 1;
    if (n4 > EE_L2C_ONE_CTA4 && k == 1) k = 2;
    int64_t s4 = (n4 + k - 1) / k;
    if (128 + (size_t)s4 * 16 * (k == 1 ? 2 : 1) > (size_t)227 * 1024) return 0;
    *K = k; *slice4 = s4; *threads = (s4 >= 1024) ? EE_L2C_MAX_THREADS : 128;
    return 1;
}
static float rms_cluster(const float *v, int64_t n, int K, int64_t slice4, int T)
{
    const int64_t n4 = n >> 2;
    float total = 0.0f;
    for (int k = 0; k < K; ++k) {
        int64_t lo4 = (int64_t)k * slice4, hi4 = lo4 + slice4 < n4 ? lo4 + slice4 : n4;
        float part[EE_L2C_MAX_THREADS];
        for (int t = 0; t < T; ++t) {
            float acc = 0.0f;
            for (int64_t i = lo4 + t; i < hi4; i += T)
                for (int c = 0; c < 4; ++c) { float e = v[4 * i + c]; acc = fmaf(e, e, acc); }
            part[t] = acc;
        }
        float wsum[32];
        for (int w = 0; w < 32; ++w) wsum[w] = 0.0f;
        for (int w = 0; w < T / 32; ++w) {
            float *p = part + 32 * w;
            for (int s = 16; s >= 1; s >>= 1)
                for (int l = 0; l < s; ++l) p[l] = p[l] + p[l + s];
            wsum[w] = p[0];
        }
        for (int s = 16; s >= 1; s >>= 1)
            for (int l = 0; l < s; ++l) wsum[l] = wsum[l] + wsum[l + s];
        total = (k == 0) ? wsum[0] : total + wsum[0];
    }
    return sqrtf(total / (float)n);
}

/* TRADES PGD-L2 step, attacks.py:391-399.  mode < 0: the order the library picks for 16-byte aligned tensors (cluster
 * kernel when the sample fits, else the three-pass kernel); mode 0: the three-pass kernel's order. */
void ee_oracle_pgd_l2_step(const float *x, const float *g, const float *x0, float *out, int B,
                           int64_t n_per, float step, float eps, int mode)
{
    int K = 0, T = 0; int64_t slice4 = 0;
    const int cluster = (mode < 0) && l2_cluster_plan(n_per, &K, &slice4, &T);
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b) {
        const float *xb = x + (int64_t)b * n_per, *gb = g + (int64_t)b * n_per, *x0b = x0 + (int64_t)b * n_per;
        float *ob = out + (int64_t)b * n_per;
        float gn = (cluster ? rms_cluster(gb, n_per, K, slice4, T) : rms_canonical(gb, n_per, NULL, 0)) + 1e-8f;      /* :391 */
        float dn;
        if (cluster) {
            for (int64_t i = 0; i < n_per; ++i) ob[i] = (xb[i] + step * (gb[i] / gn)) - x0b[i];    /* d = xa - x0, :391-394 */
            dn = rms_cluster(ob, n_per, K, slice4, T);                                                 /* :395 */
        } else {
            for (int64_t i = 0; i < n_per; ++i) ob[i] = xb[i] + step * (gb[i] / gn);               /* :391-392 */
            dn = rms_canonical(ob, n_per, x0b, 1);                                                  /* :394-395 */
        }
        int cond = dn > eps;                                       /* :396 */
        float scale = eps / dn;                                    /* :397 */
        for (int64_t i = 0; i < n_per; ++i) {
            float d = cluster ? ob[i] : ob[i] - x0b[i];
            if (cond) d = d * scale;
            ob[i] = minn(maxn(x0b[i] + d, 0.0f), 1.0f);            /* :398-399 */
        }
    }
}

/* with_gf=True blend, Tiny_ImageNet/models_tinyimagenet/resnet_EE.py:185-191: the edge map through a ZERO-padded 3x3
 * Gaussian (F.conv2d(.., padding=1)), then clamp(base + w * that, 0, 1).  Same tap order as blur3_replicate. */
static float gf_tap(const float *p, int H, int W, int i, int j, float c0, float c1, float c2)
{
    float pu = fmaf(c1, at0(p, H, W, i - 1, j), c0 * (at0(p, H, W, i - 1, j - 1) + at0(p, H, W, i - 1, j + 1)));
    float qm = fmaf(c2, at0(p, H, W, i, j), c1 * (at0(p, H, W, i, j - 1) + at0(p, H, W, i, j + 1)));
    float pd = fmaf(c1, at0(p, H, W, i + 1, j), c0 * (at0(p, H, W, i + 1, j - 1) + at0(p, H, W, i + 1, j + 1)));
    return (pu + qm) + pd;
}
void ee_oracle_gf_blend_fwd(const float *edge, const float *base, float *out, int B, int C, int H, int W,
                            float c0, float c1, float c2, float w)
{
    const size_t hw = (size_t)H * W;
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < H; ++i)
            for (int j = 0; j < W; ++j) {
                float we = w * gf_tap(edge + b * hw, H, W, i, j, c0, c1, c2);
                for (int c = 0; c < C; ++c) {
                    size_t o = ((size_t)b * C + c) * hw + (size_t)i * W + j;
                    out[o] = clamp01_nan(base[o] + we);
                }
            }
}
/* g_base = g_out * [0 <= pre <= 1]; t = sum_c w * g_base_c (fma chain); g_edge = conv^T(t) = the same symmetric taps */
int ee_oracle_gf_blend_bwd(const float *g_out, const float *edge, const float *base, float *g_edge, float *g_base,
                           int B, int C, int H, int W, float c0, float c1, float c2, float w)
{
    const size_t hw = (size_t)H * W;
    float *t = (float *)malloc(sizeof(float) * hw * (size_t)B);
    if (!t) return 1;
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b) {
        for (int i = 0; i < H; ++i)
            for (int j = 0; j < W; ++j) {
                float we = w * gf_tap(edge + b * hw, H, W, i, j, c0, c1, c2);
                float acc = 0.0f;
                for (int c = 0; c < C; ++c) {
                    size_t o = ((size_t)b * C + c) * hw + (size_t)i * W + j;
                    float pre = base[o] + we;
                    float gp = (pre >= 0.0f && pre <= 1.0f) ? g_out[o] : 0.0f;
                    acc = (c == 0) ? gp * w : fmaf(gp, w, acc);
                    if (g_base) g_base[o] = gp;
                }
                t[b * hw + (size_t)i * W + j] = acc;
            }
        if (g_edge)
            for (int i = 0; i < H; ++i)
                for (int j = 0; j < W; ++j) g_edge[b * hw + (size_t)i * W + j] = gf_tap(t + b * hw, H, W, i, j, c0, c1, c2);
    }
    free(t);
    return 0;
}

void ee_oracle_to_compare_fwd(const float *in, float *out, int64_t n, float thr)
{ for (int64_t i = 0; i < n; ++i) out[i] = to_compare(in[i], thr); }
void ee_oracle_to_compare_bwd(const float *g, const float *in, float *out, int64_t n, float thr)
{ for (int64_t i = 0; i < n; ++i) out[i] = (in[i] <= thr || in[i] > 1.001f) ? 0.0f : g[i]; }
void ee_oracle_to_eq_fwd(const float *in, float *out, int64_t n)
{ for (int64_t i = 0; i < n; ++i) out[i] = (in[i] == 0.5f) ? 1.0f : 0.0f; }
void ee_oracle_to_eq_bwd(const float *g, const float *in, float *out, int64_t n)
{ for (int64_t i = 0; i < n; ++i) out[i] = (in[i] != 0.5f) ? 0.0f : g[i]; }
void ee_oracle_safe_sign_fwd(const float *in, float *out, int64_t n)
{ for (int64_t i = 0; i < n; ++i) { float s = sgn(in[i]); out[i] = (s == 0.0f) ? -1.0f : s; } }
void ee_oracle_safe_sign_bwd(const float *g, const float *in, float *out, int64_t n)
{ for (int64_t i = 0; i < n; ++i) out[i] = (fabsf(in[i]) > 1.001f) ? 0.0f : g[i]; }

/* random start, attacks.py:15-17: x = clamp(x + noise, lo, hi) (the uniform noise is an input) */
void ee_oracle_add_clamp(const float *x, const float *noise, float *out, int64_t n, float lo, float hi)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) out[i] = minn(maxn(x[i] + noise[i], lo), hi);
}

/* AVmixup vertex + mix, attacks.py:469-471 and :476: perturb = (x - inputs)*gamma; vertex = clamp(inputs + perturb, 0, 1)
 * in fp32; the mix inputs*w + vertex*(1 - w) runs in float64 (w is a float64 tensor, torch promotes) and is cast back
 * to float at :478. */
void ee_oracle_avmixup_mix(const float *x_adv, const float *inputs, const double *weight, float *out, int B,
                           int64_t n_per, float gamma)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)B * n_per; ++i) {
        const double w = weight[i / n_per];
        const float in = inputs[i];
        const float perturb = (x_adv[i] - in) * gamma;
        const float vertex = minn(maxn(in + perturb, 0.0f), 1.0f);
        const double a = (double)in * w, b = (double)vertex * (1.0 - w);
        out[i] = (float)(a + b);
    }
}

/* ------------------------------------------------------------------------------------
 * Add_Square, utils/core.py:640-655.  The random draws are inputs:
 * stripe[B,C,W] = sign(2*rand-1) of :641, table[n_sq][2+C] = {vh, s, 2*eps*sign_c...} of :646-650.
 *   x_best = clamp(x + eps*stripe, 0, 1)                                              :641
 *   per query: x_best += new_deltas (square rows AND columns [vh, vh+s))               :648-652
 *              x_best = min(max(x_best, x - eps), x + eps); clamp(x_best, 0, 1)        :653-655
 * bwd (g == NULL: forward): multiplier of autograd through clamp (inclusive mask), torch.max /
 * torch.min (ties split 1/2 : 1/2 between the two operands, both of which depend on x).
 * ---------------------------------------------------------------------------------- */
void ee_oracle_add_square(const float *g, const float *x, const float *stripe, const float *table, float *out,
                          int B, int C, int H, int W, int n_sq, float eps)
{
#pragma omp parallel for schedule(static)
    for (int64_t plane = 0; plane < (int64_t)B * C; ++plane) {
        const int c = (int)(plane % C);
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w) {
                const int64_t i = (plane * H + h) * W + w;
                const float xv = x[i];
                const float a0 = xv + eps * stripe[plane * W + w];
                float t = minn(maxn(a0, 0.0f), 1.0f);
                float dm = (a0 >= 0.0f && a0 <= 1.0f) ? 1.0f : 0.0f;
                const float lo = xv - eps, hi = xv + eps;
                for (int q = 0; q < n_sq; ++q) {
                    const float *row = table + (size_t)q * (2 + C);
                    const int pos = (int)row[0], side = (int)row[1];
                    const int inside = (h >= pos) && (h < pos + side) && (w >= pos) && (w < pos + side);
                    const float a2 = t + (inside ? row[2 + c] : 0.0f);
                    dm = (a2 > lo) ? dm : ((a2 < lo) ? 1.0f : fmaf(0.5f, dm, 0.5f));
                    const float a3 = maxn(a2, lo);
                    dm = (a3 < hi) ? dm : ((a3 > hi) ? 1.0f : fmaf(0.5f, dm, 0.5f));
                    const float a4 = minn(a3, hi);
                    dm = (a4 >= 0.0f && a4 <= 1.0f) ? dm : 0.0f;
                    t = minn(maxn(a4, 0.0f), 1.0f);
                }
                out[i] = g ? g[i] * dm : t;
            }
    }
}

/* ------------------------------------------------------------------------------------
 * HighFreqSuppress, utils/core.py:15-55.  PARITY UNPINNED against the reference (its
 * torch.rfft / irfft calls cannot run on any torch that exists here); pinned against the torch.fft restatement
 * and against the closed form y = A x Qc^T - Bm x Qs^T (tests/test_host_logic.py).  Same five products, same
 * tables and the same fmaf chains (ascending index) as the CUDA kernel ee_hfs.cuh:
 *   T = x CB ; D = RB^T T ; G = W o D (+ the four cross terms of frequency row -r) ; V = RB G ; y = V CB^T
 * cb[N][NJp], rb[N][NIp], w[NIp][NJp]; NJ = 2r-1, NI = 2r+1, padded to multiples of 4 with zeros.
 * ---------------------------------------------------------------------------------- */
/* sum of ks partial chains in the order of an xor-butterfly seen from lane 0: ((p0+p1)+(p2+p3))+((p4+p5)+(p6+p7)) */
static float tree_sum(float *p, int ks)
{
    for (int st = 1; st < ks; st <<= 1)
        for (int i = 0; i + st < ks; i += 2 * st) p[i] = p[i] + p[i + st];
    return p[0];
}

/* ks1 / ks2: the large-plane kernel (hfs_rows_kernel) splits the K range of T = x CB over ks1 lanes and that of
 * D = RB^T T over ks2 lanes and reduces the partial chains with a butterfly; 1 / 1 for the whole-plane kernel.
 * fold: the whole-plane kernel (hfs_kernel) uses the parity of the bases along w -- cos(k th_w) even, sin(k th_w) odd about
 * w = 0 -- to halve the two large products: every row of x is folded into xe[w] = x[w] + x[N-w] (w = 1..N/2-1; w = 0 and
 * N/2 unchanged) and xo[w'] = x[w'] - x[N-w'] (w' = N/2+1..N-1); the cosine columns of T sum xe over w = 0..N/2, the sine
 * columns xo over w = N/2..N-1 (table entry exactly 0 at N/2); y[h][w] = Ye + Yo and y[h][N-w] = Ye - Yo with the cosine
 * / sine halves of the last product as two separate ascending chains. */
int ee_oracle_hfs(const float *x, float *y, const float *add, int planes, int N, int r, const float *cb,
                  const float *rb, const float *w, float gamma, int ks1, int ks2, int fold)
{
    const int NJ = 2 * r - 1, NI = 2 * r + 1;
    const int NJp = (NJ + 3) / 4 * 4, NIp = (NI + 3) / 4 * 4;
    int failed = 0;
#pragma omp parallel for schedule(static)
    for (int p = 0; p < planes; ++p) {
        float *T = (float *)malloc(sizeof(float) * ((size_t)2 * N * NJp + 2 * NIp * NJp));
        if (!T) { failed = 1; continue; }
        float *V = T + (size_t)N * NJp, *D = V + (size_t)N * NJp, *G = D + NIp * NJp;
        const float *X = x + (size_t)p * N * N;
        float *Y = y + (size_t)p * N * N;
        const int half = N / 2;
        for (int h = 0; h < N && fold; ++h) {
            float xf[1024];
            for (int q = 0; q < N; ++q) xf[q] = X[h * N + q];
            for (int q = 1; q < half; ++q) { float u = X[h * N + q], v = X[h * N + N - q]; xf[q] = u + v; xf[N - q] = v - u; }
            for (int j = 0; j < NJp; ++j) {
                float acc = 0.0f;
                if (j < r) { for (int q = 0; q <= half; ++q) acc = fmaf(xf[q], cb[q * NJp + j], acc); }
                else { for (int q = half; q < N; ++q) acc = fmaf(xf[q], cb[q * NJp + j], acc); }
                T[h * NJp + j] = acc;
            }
        }
        for (int h = 0; h < N && !fold; ++h)
            for (int j = 0; j < NJp; ++j) {
                float part[8] = {0.0f};
                for (int s = 0; s < ks1; ++s) {
                    float acc = 0.0f;
                    for (int q = s * (N / ks1); q < (s + 1) * (N / ks1); ++q) acc = fmaf(X[h * N + q], cb[q * NJp + j], acc);
                    part[s] = acc;
                }
                T[h * NJp + j] = tree_sum(part, ks1);
            }
        for (int i = 0; i < NIp; ++i)
            for (int j = 0; j < NJp; ++j) {
                float part[8] = {0.0f};
                for (int s = 0; s < ks2; ++s) {
                    float acc = 0.0f;
                    for (int h = s * (N / ks2); h < (s + 1) * (N / ks2); ++h) acc = fmaf(rb[h * NIp + i], T[h * NJp + j], acc);
                    part[s] = acc;
                }
                D[i * NJp + j] = tree_sum(part, ks2);
            }
        for (int i = 0; i < NIp; ++i)
            for (int j = 0; j < NJp; ++j) {
                float g = w[i * NJp + j] * D[i * NJp + j];
                if (j >= 1 && j < NJ) {
                    const int jcos = (j < r);
                    const int k = jcos ? j : j - (r - 1);
                    const int jc = k, js = r - 1 + k;
                    if (i == 2 * r) g = jcos ? fmaf(-gamma, D[r * NJp + js], g) : fmaf(gamma, D[r * NJp + jc], g);
                    if (i == r) g = jcos ? fmaf(gamma, D[2 * r * NJp + js], g) : fmaf(-gamma, D[2 * r * NJp + jc], g);
                }
                G[i * NJp + j] = (i < NI && j < NJ) ? g : 0.0f;
            }
        for (int h = 0; h < N; ++h)
            for (int j = 0; j < NJp; ++j) {
                float acc = 0.0f;
                for (int i = 0; i < NI; ++i) acc = fmaf(rb[h * NIp + i], G[i * NJp + j], acc);
                V[h * NJp + j] = acc;
            }
        for (int h = 0; h < N && fold; ++h)
            for (int q = 0; q <= half; ++q) {
                float ye = 0.0f, yo = 0.0f;
                for (int j = 0; j < r; ++j) ye = fmaf(V[h * NJp + j], cb[q * NJp + j], ye);
                for (int j = r; j < NJp; ++j) yo = fmaf(V[h * NJp + j], cb[q * NJp + j], yo);
                const size_t o1 = (size_t)h * N + q, o2 = (size_t)h * N + N - q, pb = (size_t)p * N * N;
                const float y1 = ye + yo, y2 = ye - yo;
                Y[o1] = add ? y1 + add[pb + o1] : y1;                                   /* y = H x + add */
                if (q != 0 && q != half) Y[o2] = add ? y2 + add[pb + o2] : y2;
            }
        for (int h = 0; h < N && !fold; ++h)
            for (int q = 0; q < N; ++q) {
                float acc = 0.0f;
                for (int j = 0; j < NJp; ++j) acc = fmaf(V[h * NJp + j], cb[q * NJp + j], acc);
                Y[h * N + q] = add ? acc + add[(size_t)p * N * N + h * N + q] : acc;    /* y = H x + add */
            }
        free(T);
    }
    return failed;
}

int ee_oracle_version(void) { return 1; }

/* number of host threads the parallel loops above will use (reported as `cores` by bench.py) */
#ifdef _OPENMP
#include <omp.h>
int ee_oracle_threads(void) { return omp_get_max_threads(); }
#else
int ee_oracle_threads(void) { return 1; }
#endif
