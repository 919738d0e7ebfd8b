/*
 * edge_b200.h -- C ABI of libedge_b200.so: the B200 (sm_100a) implementation of the
 * edge-enhancement transform and the PGD/FGSM inner-loop updates of Aiqz/Edge-Enhancement.
 *
 * This is the drop-in boundary.  The reference has no native code; what a maintainer binds
 * instead of is a chain of PyTorch eager ops.  Each entry point names the reference lines it
 * replaces (paths relative to the reference repo).  INTEGRATION.md shows the ctypes stub.
 *
 * Contract (all entry points)
 *   - plain pointers and sizes; no torch / C++ types; no exceptions cross the boundary;
 *   - every pointer is a DEVICE pointer to dense fp32 data in EEParams.layout (NCHW by default), owned by
 *     the caller; the library never allocates, frees or retains device memory and keeps no
 *     global device state, so every call is CUDA-graph capturable;
 *   - work is enqueued asynchronously on `stream` (a cudaStream_t passed as void*) of the
 *     CURRENT device; no host synchronisation; re-entrant and thread-safe (DataParallel worker
 *     threads, autograd engine threads);
 *   - returns EE_OK (0) or a negative EE_ERR_* / positive cudaError_t; ee_last_error() gives a
 *     thread-local message.  In-place use is allowed only where stated.
 */
#ifndef EDGE_B200_H_
#define EDGE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EE_VERSION 201 /* 0.2.1 */

enum {
    EE_OK = 0,
    EE_ERR_INVALID_ARG = -1, /* null pointer, non-positive size, bad enum             */
    EE_ERR_UNSUPPORTED = -2, /* valid request this build does not implement            */
    EE_ERR_TOO_LARGE = -3    /* a single image row strip does not fit in shared memory */
};

/* filter variants == the three reference modules */
enum {
    EE_VARIANT_STEP125 = 0, /* utils/core.py:509-585  CannyFilter_step125_1 */
    EE_VARIANT_CANNY = 1,   /* utils/core.py:148-326  CannyFilter           */
    EE_VARIANT_BPDA = 2     /* utils/core.py:386-505  CannyFilter_BPDA      */
};

/* memory layout of the [B,C,H,W] image tensors (the edge map [B,1,H,W] is the same in both).
 * EE_LAYOUT_NHWC = torch.channels_last; implemented by the fused blend entry points for C == 3,
 * W % 4 == 0 (EE_ERR_UNSUPPORTED otherwise -- the caller converts). */
enum { EE_LAYOUT_NCHW = 0, EE_LAYOUT_NHWC = 1 };

/* Filter description, passed by value into kernel-parameter space (no device constants).
 * Mirrors the module state + forward() arguments of the reference filters. */
typedef struct EEParams {
    int32_t variant;    /* EE_VARIANT_*                                                        */
    int32_t layout;     /* EE_LAYOUT_NCHW or EE_LAYOUT_NHWC                                    */
    float gauss[9];     /* weight_gaussian, row major (core.py:163-165).  Must have the         */
                        /* corner/edge/centre symmetry every get_gaussian_kernel(3,mu,sigma) has */
    float sobel[9];     /* weight_sobel_x, row major (core.py:175-178); must equal              */
                        /* get_sobel_kernel(3); sobel_y is its transpose (core.py:180)          */
    float alpha;        /* magnitude gate (core.py:264, :575); ignored by BPDA                  */
    float low_thr;      /* forward(low_threshold=...), already /255 (resnet_EE.py:127)          */
    float high_thr;     /* forward(high_threshold=...)                                          */
    int32_t has_low;    /* low_threshold is not None                                            */
    int32_t has_high;   /* high_threshold is not None (STEP125 requires it, core.py:578-583)    */
    int32_t hysteresis; /* forward(hysteresis=...); ignored by STEP125                          */
    int32_t flags;      /* EE_FLAG_* bits, 0 by default                                         */
} EEParams;

/* EEParams.flags.
 * EE_FLAG_NAN_COMPAT: reproduce the reference backward's NaNs.  autograd differentiates (gx^2 + gy^2) ** 0.5
 * (core.py:250, :453, :571) at magnitude 0 as 0 * inf, so the reference's input gradient is NaN on the 5 x 5
 * neighbourhood of every pixel whose gradient magnitude is exactly 0 (flat regions: MNIST backgrounds, saturated
 * patches) and torch.sign(NaN) = 0 freezes those pixels in PGD / FGSM.  Default (flag clear): the sub-gradient 0.
 * With the flag the backward entry points write NaN into dL/dSgx, dL/dSgy there and let it spread through the two
 * adjoint stencils exactly like the reference; they then run the shape-generic kernels (NCHW only). */
enum { EE_FLAG_NAN_COMPAT = 1 };

/* ---- edge filter: module-level drop-in -------------------------------------------------- */

/* edge[B,1,H,W] = CannyFilter*.forward(x[B,C,H,W], low, high, hysteresis)
 * replaces utils/core.py:222-326, :426-505, :549-585. */
int ee_edge_fwd_f32(const float* x, float* edge, int B, int C, int H, int W,
                    const EEParams* p, void* stream);

/* g_x[B,C,H,W] = d(edge)/d(x)^T g_edge[B,1,H,W]; replaces autograd through the above incl.
 * To_compare/To_eq/BinaryConnectDeterministic.backward (core.py:138-145, :350-358, :375-382).
 * Forward intermediates are recomputed from x (no saved tensors).  Sub-gradient of
 * sqrt at 0 is 0 unless EE_FLAG_NAN_COMPAT asks for the reference's NaNs. */
int ee_edge_bwd_f32(const float* g_edge, const float* x, float* g_x, int B, int C, int H, int W,
                    const EEParams* p, void* stream);

/* ---- edge filter fused with the blend: model-level drop-in ------------------------------- */

/* out[B,C,H,W] = clamp(base + w * edge(x), 0, 1); optionally also writes edge[B,1,H,W].
 * replaces core.py filter forward + e.g. Tiny_ImageNet/models_tinyimagenet/resnet_EE.py:182-191
 * (gf=False).  One pass over HBM: reads x and base, writes out. */
int ee_edge_blend_fwd_f32(const float* x, const float* base, float* out, float* edge_or_null,
                          int B, int C, int H, int W, const EEParams* p, float w, void* stream);

/* Adjoint of the above in one pass: reads g_out, x, base; writes
 *   g_x    (edge path only; the caller's autograd adds HFS^T(g_base)), may be NULL
 *   g_base = g_out * [0 <= base + w*edge <= 1],                        may be NULL */
int ee_edge_blend_bwd_f32(const float* g_out, const float* x, const float* base, float* g_x_or_null,
                          float* g_base_or_null, int B, int C, int H, int W, const EEParams* p,
                          float w, void* stream);

/* The `gf` option of the *_EE models (with_gf=True: Tiny_ImageNet/models_tinyimagenet/resnet_EE.py:185-187 and its
 * copies): the edge map goes through a ZERO-padded 3 x 3 Gaussian (get_gaussian_kernel(3, 0, 1)) before the blend,
 *     out = clamp(base + w * conv2d(edge, gauss, padding=1), 0, 1).
 * No reference config enables it, so it is not fused into the filter kernels: the caller runs ee_edge_fwd_f32 first and
 * ee_edge_bwd_f32 last.  `gauss` is a HOST pointer to the nine taps (they travel in kernel-parameter space like
 * EEParams.gauss); every other pointer is a device pointer.  fwd reads edge[B,1,H,W] and base, writes out.  bwd reads g_out, edge, base and writes
 * g_edge[B,1,H,W] = conv2d^T(w * sum_c g_pre_c) (may be NULL) and g_base = g_pre = g_out * [0 <= pre <= 1] (may be NULL). */
int ee_gf_blend_fwd_f32(const float* edge, const float* base, float* out, int B, int C, int H, int W,
                        const float gauss[9], float w, void* stream);
int ee_gf_blend_bwd_f32(const float* g_out, const float* edge, const float* base, float* g_edge_or_null,
                        float* g_base_or_null, int B, int C, int H, int W, const float gauss[9], float w, void* stream);

/* ---- strided variants (`const int64_t xs[4]`) ------------------------------------------
 * Element strides of a [B,C,H,W] tensor in torch.Tensor.stride() order; a NULL EEStrides pointer means "dense in
 * EEParams.layout".  Accepted without a copy: dense NCHW, dense NHWC (= torch.channels_last, subject to the NHWC rules
 * above), and ANY tensor whose column stride is 1 -- sliced batches (x[::2], x[lo:hi]), channel slices (x[:, 1:4]),
 * expanded batches (stride 0) and spatial crops (x[:, :, 16:240, 16:240]).  Dense tensors run the tuned kernels; other
 * strides run the shape-generic kernels on the tensor in place (128-bit path when every stride and base address is a
 * multiple of 4 floats), which costs less than the contiguous copy it replaces.  A column stride other than 1 (outside
 * channels_last) returns EE_ERR_UNSUPPORTED.  Outputs may be strided too; tensors must not overlap. */
typedef struct EEStrides { int64_t n, c, h, w; } EEStrides;

int ee_edge_fwd_strided_f32(const float* x, const EEStrides* xs, float* edge, const EEStrides* es, int B, int C, int H, int W,
                            const EEParams* p, void* stream);
int ee_edge_bwd_strided_f32(const float* g_edge, const EEStrides* ges, const float* x, const EEStrides* xs, float* g_x,
                            const EEStrides* gxs, int B, int C, int H, int W, const EEParams* p, void* stream);
int ee_edge_blend_fwd_strided_f32(const float* x, const EEStrides* xs, const float* base, const EEStrides* bs, float* out,
                                  const EEStrides* os, float* edge_or_null, const EEStrides* es, int B, int C, int H, int W,
                                  const EEParams* p, float w, void* stream);
int ee_edge_blend_bwd_strided_f32(const float* g_out, const EEStrides* gs, const float* x, const EEStrides* xs, const float* base,
                                  const EEStrides* bs, float* g_x_or_null, const EEStrides* gxs, float* g_base_or_null,
                                  const EEStrides* gbs, int B, int C, int H, int W, const EEParams* p, float w, void* stream);

/* Workspace the edge entry points need from the caller: always 0 (recompute formulation). */
size_t ee_aux_bytes(int B, int C, int H, int W, int variant);

/* ---- attack inner-loop updates (elementwise, n = number of floats; out may alias x) ------ */

/* out = clamp(min(max(x + alpha_signed*sign(g), x0-eps), x0+eps), lo, hi)
 * replaces utils/attacks.py:25-27 and its copies :52-54 :82-84 :257-259 :298-300 :318-320
 * :353-355 :414-416 :466-468 :505-507 (targeted ones pass alpha_signed = -step_size). */
int ee_pgd_linf_step_f32(const float* x, const float* g, const float* x0, float* out, int64_t n,
                         float alpha_signed, float eps, float lo, float hi, void* stream);

/* out = clamp(x + alpha_signed*sign(g), lo, hi); replaces utils/attacks.py:121-126. */
int ee_fgsm_step_f32(const float* x, const float* g, float* out, int64_t n, float alpha_signed,
                     float lo, float hi, void* stream);

/* delta = clamp(delta + alpha*sign(g), -eps, eps) in place; x_adv = clamp(x0 + delta, lo, hi)
 * (x_adv may be NULL).  replaces ImageNet/free_imagenet/AT_hfs_canny_free_imagenet_ddp.py
 * :330-332 + :314-315 and ImageNet/fgsm_imagenet/main_fast.py:233-235,:246-253. */
int ee_free_at_step_f32(float* delta, const float* g, const float* x0, float* x_adv_or_null,
                        int64_t n, float alpha, float eps, float lo, float hi, void* stream);

/* CW L-inf inner update, replaces utils/attacks.py:213-222:
 * t = adv + step*sign(g); t = max(min(t, x+magnitude), x-magnitude); t = clamp(t,0,1);
 * out = max(min(t, max_x), min_x). */
int ee_cw_linf_step_f32(const float* adv, const float* g, const float* x, const float* min_x,
                        const float* max_x, float* out, int64_t n, float step, float magnitude,
                        void* stream);

/* out = clamp(x + noise, lo, hi): the random start of every attack, replaces utils/attacks.py:15-17 (:45-47, :74-77,
 * :454-456, :495-497); the noise itself stays a torch uniform_ draw so that a seeded run follows the reference. */
int ee_add_clamp_f32(const float* x, const float* noise, float* out, int64_t n, float lo, float hi, void* stream);

/* AVmixup vertex + mix, replaces utils/attacks.py:469-471 + :476 (and :508-510 + :515):
 * vertex = clamp(inputs + (x_adv - inputs)*gamma, 0, 1) (fp32); out = float(inputs*w_b + vertex*(1 - w_b)) evaluated in
 * double, w = the reference's float64 np.random.beta weights, one per sample (device pointer, [B]). */
int ee_avmixup_mix_f32(const float* x_adv, const float* inputs, const double* weight, float* out, int B,
                       int64_t n_per_sample, float gamma, void* stream);

/* TRADES PGD-L2 step with per-sample RMS norms, replaces utils/attacks.py:391-399 and
 * l2_norm (:360-366).  One CTA per sample; out must NOT alias x. */
int ee_pgd_l2_step_f32(const float* x, const float* g, const float* x0, float* out, int B,
                       int64_t n_per_sample, float step, float eps, void* stream);

/* One whole iteration of the edge-enhanced PGD hot path (the loop body of utils/attacks.py:19-27 around a model whose
 * front end is the fused edge filter + blend) enqueued by ONE call, for callers that hold the upstream gradient g_out
 * already (benchmarks, CUDA-graph capture, custom training loops):
 *     out    = clamp(base + w*edge(x), 0, 1)                      (skipped when out_or_null == NULL)
 *     g_x, g_base = adjoint of the above applied to g_out
 *     x_next = clamp(min(max(x + alpha_signed*sign(g_x), x0-eps), x0+eps), 0, 1)
 * Same kernels, same arithmetic as the three separate entry points; x_next must not alias x. */
int ee_edge_pgd_iteration_f32(const float* x, const float* base, const float* g_out, const float* x0,
                              float* out_or_null, float* g_x, float* g_base_or_null, float* x_next,
                              int B, int C, int H, int W, const EEParams* p, float w,
                              float alpha_signed, float eps, void* stream);

/* ---- straight-through helper Functions (elementwise; out may alias g/in) ----------------- */
int ee_to_compare_fwd_f32(const float* in, float* out, int64_t n, float thr, void* stream);                 /* core.py:338-347 */
int ee_to_compare_bwd_f32(const float* g, const float* in, float* out, int64_t n, float thr, void* stream); /* core.py:350-358 */
int ee_to_eq_fwd_f32(const float* in, float* out, int64_t n, void* stream);                                 /* core.py:364-372 */
int ee_to_eq_bwd_f32(const float* g, const float* in, float* out, int64_t n, void* stream);                 /* core.py:375-382 */
int ee_safe_sign_fwd_f32(const float* in, float* out, int64_t n, void* stream);                             /* core.py:115-118,:130-135 */
int ee_safe_sign_bwd_f32(const float* g, const float* in, float* out, int64_t n, void* stream);             /* core.py:138-145 */

/* ---- Add_Square: the random stripe + square perturbation of the *_square models ---- */

/* out[B,C,H,W] = Add_Square.forward(x), replaces utils/core.py:640-655.  The random draws stay with the caller
 * (torch CPU generator, same calls and order as the reference): stripe[B,C,W] = sign(2*rand-1) (core.py:641),
 * table[n_sq][2+C] floats = {vh, s, 2*eps*sign_c ...} per query (core.py:646-650; the square is rows AND columns
 * [vh, vh+s)).  t = clamp(x + eps*stripe, 0, 1); per query t = clamp(min(max(t + d, x-eps), x+eps), 0, 1). */
int ee_add_square_fwd_f32(const float* x, const float* stripe, const float* table, float* out,
                          int B, int C, int H, int W, int n_sq, float eps, void* stream);

/* g_x = g * d(out)/d(x) of the above (autograd through clamp / torch.max / torch.min incl. the even split on
 * ties), recomputed from x: one pass, no saved tensors. */
int ee_add_square_bwd_f32(const float* g, const float* x, const float* stripe, const float* table, float* g_x,
                          int B, int C, int H, int W, int n_sq, float eps, void* stream);

/* ---- HighFreqSuppress: the square low-pass in front of the *_EE models ---------------- */

/* y[planes][N][N] = HighFreqSuppress(N, N, r)(x), replaces utils/core.py:47-52 (rfft -> mask -> irfft; the reference's
 * torch.rfft no longer exists, so parity is pinned to the torch.fft restatement only).  The operator is the real
 * symmetric  y = A x Qc^T - Bm x Qs^T ; it is evaluated as five small dense products per plane on the caller-supplied
 * tables cb[N][NJp] (1, cos k, sin k for k < r; NJp = 2r-1 rounded up to 4), rb[N][NIp] (1, cos k, sin k for k <= r),
 * w[NIp][NJp] = alpha_i * beta_j and gamma = 2/N^2 (core.HighFreqSuppress builds them).  Self-adjoint: the backward is the
 * same call on the upstream gradient; with add_or_null != NULL the result is y = H x + add (elementwise, may alias y), which
 * lets the backward of the *_EE front end accumulate H(g_base) into the edge-path gradient without an extra pass.
 * y must not alias x.  ee_hfs_supported(N, r) tells whether a kernel exists. */
int ee_hfs_f32(const float* x, float* y, const float* add_or_null, int planes, int N, int r, const float* cb,
               const float* rb, const float* w, float gamma, void* stream);
int ee_hfs_supported(int N, int r);

/* The same operator with the four dense products on the tensor cores (tcgen05.mma kind::tf32, 3 x TF32 split, accumulators in
 * tensor memory; csrc/ee_hfs_tc.cuh).  Same arguments and tables as ee_hfs_f32; 64 x 64 planes with radius 8 only
 * (ee_hfs_tc_supported), EE_ERR_UNSUPPORTED otherwise.  NOT bit-identical to ee_hfs_f32 / the oracle: the accumulation order
 * inside the tensor core is unspecified; max abs error 1.3e-6 against float64 on [0,1] inputs (ee_hfs_f32: 0.6e-6).  1.5x
 * faster than ee_hfs_f32 (103 vs 153 us at 4096x3x64x64), which stays the default because it is exact.
 * x and y go through TMA tensor copies (the call encodes two tensor maps on the host); add_or_null == y is the fast
 * accumulate path (TMA reduction store). */
int ee_hfs_tc_f32(const float* x, float* y, const float* add_or_null, int planes, int N, int r, const float* cb,
                  const float* rb, const float* w, float gamma, void* stream);
int ee_hfs_tc_supported(int N, int r);

/* ---- misc -------------------------------------------------------------------------------- */
const char* ee_last_error(void); /* thread-local, never NULL */
int ee_version(void);            /* EE_VERSION */

/* Tuning knob for benchmarks/tests: force the row-strip height of the tiled edge kernels
 * (0 = heuristic) and the staging path (0 = auto, 1 = generic kernels only, 4 = tuned kernels even for wide images,
 * 3 = strip kernels instead of chunk-aligned tiles for wide images, 5 = experimental thread-block-cluster backward
 * for 3x224x224, 6 = always stage the x tiles of wide images by TMA tensor copies, 7 = never use TMA staging).
 * Process-wide; returns EE_OK.  Not needed for normal use. */
int ee_set_tuning(int strip_rows_fwd, int strip_rows_bwd, int staging);

#ifdef __cplusplus
}
#endif
#endif /* EDGE_B200_H_ */
