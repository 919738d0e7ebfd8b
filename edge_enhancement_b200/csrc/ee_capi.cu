// ee_capi.cu -- the C ABI of libedge_b200.so (include/edge_b200.h): argument validation, launch
// configuration and error reporting.  No torch types, no allocation, no host synchronisation.
#include "../../include/edge_b200.h"
#include "ee_attack.cuh"
#include "ee_edge_canny.cuh"
#include "ee_edge_canny_fast.cuh"
#include "ee_edge_cluster.cuh"
#include "ee_edge_tiles.cuh"
#include "ee_edge_canny_tiles.cuh"
#include "ee_edge_stream.cuh"
#include "ee_edge_fast.cuh"
#include "ee_edge_step125.cuh"
#include "ee_square.cuh"
#include "ee_hfs.cuh"
#include "ee_hfs_tc.cuh"
#include "ee_gf.cuh"
#include "ee_pgd_l2.cuh"

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

// The library is compiled as six translation units from this one source (edge_enhancement_b200/_build.py, in
// parallel): -DEE_PART=1 the C ABI + elementwise kernels, 2 / 3 the step125 forward / backward kernel families,
// 4 / 5 the Canny + BPDA forward / backward families, 6 the row-streaming Canny + BPDA kernels for wide images.
// EE_PART undefined (0) builds everything in one unit.
#ifndef EE_PART
#define EE_PART 0
#endif
#define EE_HAS(k) (EE_PART == 0 || EE_PART == (k))

namespace ee_shared {       // process-wide state, defined in part 1
extern thread_local char g_err[512];
extern std::atomic<int> g_th_fwd, g_th_bwd, g_staging;
#if EE_HAS(1)
thread_local char g_err[512] = "";
std::atomic<int> g_th_fwd{0}, g_th_bwd{0}, g_staging{0};
#endif
// kernel-family dispatchers (one translation unit each)
int fwd_step125(ee::EdgeArgs& a, const EEParams* p, int B, bool blend, bool vec_ok, bool nhwc, cudaStream_t s);
int bwd_step125(ee::EdgeArgs& a, const EEParams* p, int B, bool blend, bool vec_ok, bool nhwc, cudaStream_t s);
int fwd_canny(ee::EdgeArgs& a, const EEParams* p, int B, bool blend, bool vec_ok, bool nhwc, cudaStream_t s);
int bwd_canny(ee::EdgeArgs& a, const EEParams* p, int B, bool blend, bool vec_ok, bool nhwc, cudaStream_t s);
int canny_stream(ee::EdgeArgs& a, const EEParams* p, int B, bool blend, bool bwd, cudaStream_t s);
}  // namespace ee_shared

namespace {

using ee_shared::g_err;
using ee_shared::g_th_fwd;
using ee_shared::g_th_bwd;
using ee_shared::g_staging;

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
int cuda_fail(cudaError_t e, const char* what) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return (int)e;
}
bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

constexpr int kMaxSmem = 227 * 1024;
constexpr int kThreads = 256;
constexpr int kMaxDevices = 32;

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a driver round trip (~2 us); a kernel needs it once per
// device, not once per launch.  The largest value set so far is remembered per (kernel address, device).
struct SmemOnce { std::atomic<int> have[kMaxDevices]; };
inline int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);             // thread-local lookup in the runtime, no driver call
    return dev;
}
template <typename K>
int ensure_smem(SmemOnce& once, K kernel, size_t smem) {
    if (smem <= 48 * 1024) return EE_OK;
    const int dev = current_device();
    if (dev >= 0 && dev < kMaxDevices && once.have[dev].load(std::memory_order_relaxed) >= (int)smem) return EE_OK;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
    if (dev >= 0 && dev < kMaxDevices) once.have[dev].store((int)smem, std::memory_order_relaxed);
    return EE_OK;
}
// table of SmemOnce slots keyed by the kernel's address (lock-free, fixed size; a full table just re-sets the attribute)
inline SmemOnce& smem_slot(const void* fn) {
    constexpr int kSlots = 512;
    static std::atomic<const void*> keys[kSlots];
    static SmemOnce slots[kSlots];
    static SmemOnce overflow;
    size_t h = (reinterpret_cast<uintptr_t>(fn) >> 4) * 0x9E3779B97F4A7C15ull >> 40;
    for (int probe = 0; probe < kSlots; ++probe) {
        const int i = (int)((h + probe) % kSlots);
        const void* k = keys[i].load(std::memory_order_acquire);
        if (k == fn) return slots[i];
        if (k == nullptr) {
            const void* expect = nullptr;
            if (keys[i].compare_exchange_strong(expect, fn, std::memory_order_acq_rel) || expect == fn) return slots[i];
        }
    }
    for (int d = 0; d < kMaxDevices; ++d) overflow.have[d].store(0, std::memory_order_relaxed);
    return overflow;
}
template <typename K>
int ensure_smem(K kernel, size_t smem) {
    if (smem <= 48 * 1024) return EE_OK;
    return ensure_smem(smem_slot(reinterpret_cast<const void*>(kernel)), kernel, smem);
}
inline int sm_count() {
    static std::atomic<int> cached[kMaxDevices];
    const int dev = current_device();
    if (dev >= 0 && dev < kMaxDevices) {
        const int c = cached[dev].load(std::memory_order_relaxed);
        if (c > 0) return c;
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (dev >= 0 && dev < kMaxDevices) cached[dev].store(sms, std::memory_order_relaxed);
    return sms;
}

// Validate EEParams and extract the three Gaussian taps.
int check_params(const EEParams* p, float& c0, float& c1, float& c2) {
    if (!p) return fail(EE_ERR_INVALID_ARG, "EEParams is null");
    if (p->variant < EE_VARIANT_STEP125 || p->variant > EE_VARIANT_BPDA)
        return fail(EE_ERR_INVALID_ARG, "unknown variant %d", p->variant);
    if (p->layout != EE_LAYOUT_NCHW && p->layout != EE_LAYOUT_NHWC) return fail(EE_ERR_INVALID_ARG, "unknown layout %d", p->layout);
    if ((p->flags & ~EE_FLAG_NAN_COMPAT) != 0) return fail(EE_ERR_INVALID_ARG, "EEParams.flags has unknown bits (0x%x)", p->flags);
    const float* g = p->gauss;
    if (!(g[0] == g[2] && g[0] == g[6] && g[0] == g[8] && g[1] == g[3] && g[1] == g[5] && g[1] == g[7]))
        return fail(EE_ERR_UNSUPPORTED, "gauss[9] lacks the corner/edge/centre symmetry of get_gaussian_kernel(3,..)");
    static const float kSobel[9] = {-0.5f, 0.f, 0.5f, -1.f, 0.f, 1.f, -0.5f, 0.f, 0.5f};
    for (int i = 0; i < 9; ++i)
        if (p->sobel[i] != kSobel[i]) return fail(EE_ERR_UNSUPPORTED, "sobel[9] must equal get_sobel_kernel(3)");
    if (p->variant == EE_VARIANT_STEP125 && !p->has_high)
        return fail(EE_ERR_INVALID_ARG, "CannyFilter_step125_1 needs high_threshold (reference raises UnboundLocalError, core.py:578-583)");
    c0 = g[0]; c1 = g[1]; c2 = g[4];
    return EE_OK;
}

struct Launch { int vec, threads, GX, RY, TH, tiles; size_t smem; int TW, tiles_x, planeW, halo; int even_planes = 0; };

// ---- exact threshold cut-offs in u = mag^2 space (sqrt_rn is monotonic) ----------------------
float f_of(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }
// largest fp32 u >= 0 with pred(u) true, where pred is true at 0 and monotonically turns false
template <typename P>
float last_true(P pred) {
    uint32_t lo = 0, hi = 0x7f800000u;          // +0 .. +inf
    if (pred(f_of(hi))) return f_of(hi);
    while (hi - lo > 1) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if (pred(f_of(mid))) lo = mid; else hi = mid;
    }
    return f_of(lo);
}
// mag > thr  <=>  u > cut
float cut_gt(float thr) {
    if (thr < 0.0f) return -1.0f;
    return last_true([thr](float u) { return sqrtf(u) <= thr; });
}
// mag < alpha  <=>  u < cut
float cut_lt(float alpha) {
    if (!(alpha > 0.0f)) return 0.0f;
    const float m = last_true([alpha](float u) { return sqrtf(u) < alpha; });
    return nextafterf(m, INFINITY);
}

// Launch plan of the tuned (fast) kernels: tiles of TH rows x TW columns; planes are (TW + 8 halo) columns
// wide with 8 floats of padding per row; R rows per thread chunk.  Images up to 128 columns use one
// column tile (full-width strips); wider ones are cut into ~64-column tiles.
int plan_fast(int H, int W, int R, int rows_per_th, int rows_fixed, int max_halo_rows, int budget_bytes, int forced_th,
              Launch& L, int halo_cols = 4, int max_full_width = 128, int pref_th_tiled = 0) {
    L.vec = 4;
    L.halo = halo_cols;
    // Full-width strips up to max_full_width columns (measured at 224 px: the forward is faster with
    // full-width strips of ~18 rows, 6.3 vs 5.4 TB/s; the backward with 56+8-column tiles, 3.9-4.3 vs 3.4 TB/s)
    int tw_target = 64;
#ifdef EE_TUNING_ENV        // tuning experiments only (-DEE_TUNING_ENV): never read the environment in a production build
    if (const char* e = getenv("EE_TILE_COLS")) { const int v = atoi(e); if (v >= 8) { tw_target = v; max_full_width = 0; } }
#endif
    if (W <= max_full_width || tw_target >= W) { L.tiles_x = 1; L.TW = W; L.planeW = W; }
    else {
        L.tiles_x = (W + tw_target - 1) / tw_target;
        int tw = (W + L.tiles_x - 1) / L.tiles_x;
        L.TW = (tw + 3) & ~3;
        L.tiles_x = (W + L.TW - 1) / L.TW;
        L.planeW = L.TW + 2 * halo_cols;
    }
    const int G = L.planeW / 4;
    const size_t row_bytes = (size_t)(L.planeW + ee::kPadW) * sizeof(float);
    // the kernels' regions are min(TH + k, H) rows each (halo rows outside the image do not exist); the k's
    // are encoded as rows_per_th regions whose halos add up to rows_fixed: step125 fwd {4,2}, bwd {8,6,4},
    // Canny fwd {8,6,4}, Canny bwd {12,10,8,4,4}
    auto smem_of = [&](int th) {
        int halos[5] = {0, 0, 0, 0, 0};
        if (rows_per_th == 2) { halos[0] = 4; halos[1] = 2; }
        else if (rows_per_th == 3) { halos[0] = 8; halos[1] = 6; halos[2] = 4; }
        else { halos[0] = 12; halos[1] = 10; halos[2] = 8; halos[3] = 4; halos[4] = 4; }
        size_t rows = 0;
        for (int i = 0; i < rows_per_th; ++i) rows += (size_t)(th + halos[i] < H ? th + halos[i] : H);
        return rows * row_bytes;
    };
    int th;
    if (forced_th > 0) {
        th = forced_th < H ? forced_th : H;
    } else {
        long fit = ((long)(budget_bytes / row_bytes) - rows_fixed) / rows_per_th;
        th = (int)(fit < 4 ? 4 : (fit > H ? H : fit));
        const int tiles = (H + th - 1) / th;
        th = (H + tiles - 1) / tiles;
        // measured preference for column-tiled images (Canny backward at 224 px: 48-row strips 2.06 TB/s, the
        // equalised 56-row strips 1.49 TB/s)
        if (L.tiles_x > 1 && pref_th_tiled > 0 && pref_th_tiled <= fit && pref_th_tiled < H) th = pref_th_tiled;
        // strips that are a multiple of the R = 4 rows a thread slides over keep every chunk full
        // (measured at 224 px: 19-row strips 5.8 TB/s, 16- or 28-row strips 6.3 TB/s)
        if (L.tiles_x == 1 && th < H && (th & 3)) {
            const int up = (th + 3) & ~3;
            th = (up <= fit || th < 4) ? up : (th & ~3);
            if (th > H) th = H;
        }
    }
    while (th > 1 && smem_of(th) > (size_t)kMaxSmem) --th;
    if (smem_of(th) > (size_t)kMaxSmem) return fail(EE_ERR_TOO_LARGE, "tile of %d columns does not fit in shared memory", L.planeW);
    L.TH = th;
    L.tiles = ((H + th - 1) / th) * L.tiles_x;
    L.smem = smem_of(th);
    const int max_chunks = (th + max_halo_rows + R - 1) / R;
    if (G >= kThreads) { L.GX = kThreads; L.RY = 1; }
    else {
        L.GX = G;
        int ry = kThreads / G;
        if (ry > max_chunks) ry = max_chunks;
        if (ry < 1) ry = 1;
        L.RY = ry;
    }
    L.threads = ((L.GX * L.RY + 31) / 32) * 32;
    return EE_OK;
}

template <typename K>
int launch_fast(K kernel, const Launch& L, int B, const ee::FastArgs& a, cudaStream_t s, const char* name) {
    if (int rc = ensure_smem(kernel, L.smem)) return rc;
    const long long grid = (long long)B * L.tiles;
    if (grid > 0x7fffffffLL) return fail(EE_ERR_TOO_LARGE, "too many tiles");
    kernel<<<(unsigned)grid, L.threads, L.smem, s>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, name);
    return EE_OK;
}

// whole-image tiles whose height is a compile-time multiple of R = 4: one chunk per thread (HT instantiations)
template <typename K>
int launch_fast_even(K kernel, Launch L, int B, const ee::FastArgs& a, cudaStream_t s, const char* name) {
    L.RY = a.e.H / 4;
    L.threads = ((L.GX * L.RY + 31) / 32) * 32;
    const int planes = L.even_planes;
    if (planes) L.smem = (size_t)planes * a.e.H * a.Wp * sizeof(float);     // whole-image planes the kernel really uses
    return launch_fast(kernel, L, B, a, s, name);
}
bool whole_image(const Launch& L, const ee::FastArgs& f, int side) {
    return L.tiles_x == 1 && L.TH == f.e.H && f.e.H == side && f.e.W == side && L.GX == side / 4;
}

void fill_fast(ee::FastArgs& f, const ee::EdgeArgs& a, const Launch& L);

// One thread-block cluster per image (ee_edge_cluster.cuh): CS CTAs x TH rows, halo rows through DSMEM.
template <typename K>
int launch_cluster(K kernel, int W, int TH, int CS, int B, const ee::EdgeArgs& e, cudaStream_t s, const char* name) {
    Launch L;
    L.vec = 4; L.TH = TH; L.tiles = CS; L.GX = W / 4; L.RY = TH / 4; L.TW = W; L.tiles_x = 1; L.planeW = W; L.halo = 4;
    L.threads = ((L.GX * L.RY + 31) / 32) * 32;
    if (L.threads < 256) L.threads = 256;
    L.smem = (size_t)3 * TH * (W + ee::kPadW) * sizeof(float);
    ee::FastArgs f;
    fill_fast(f, e, L);
    if (int rc = ensure_smem(kernel, L.smem)) return rc;
    cudaError_t err;
    if ((long long)B * CS > 0x7fffffffLL) return fail(EE_ERR_TOO_LARGE, "too many tiles");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * CS));
    cfg.blockDim = dim3((unsigned)L.threads);
    cfg.dynamicSmemBytes = L.smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    err = cudaLaunchKernelEx(&cfg, kernel, f);
    if (err != cudaSuccess) return cuda_fail(err, name);
    return EE_OK;
}

bool fast_eligible(const ee::EdgeArgs& a, bool vec_ok);

// Chunk-aligned tiles for wide images (ee_edge_tiles.cuh): <= 56 x 56 outputs per CTA, 64 x 64 planes, 16 x 16 chunks.
void plan_tiles(int H, int W, Launch& L, int max_tile = 56, int planes = 3) {
    int ty = (H + max_tile - 1) / max_tile, tx = (W + max_tile - 1) / max_tile;
    L.TH = (((H + ty - 1) / ty) + 3) & ~3;
    L.TW = (((W + tx - 1) / tx) + 3) & ~3;
    ty = (H + L.TH - 1) / L.TH;
    L.tiles_x = (W + L.TW - 1) / L.TW;
    L.tiles = ty * L.tiles_x;
    L.vec = 4; L.planeW = 64; L.halo = 4; L.GX = 16; L.RY = 16; L.threads = 256;
    L.smem = (size_t)planes * (L.TH + (64 - max_tile)) * (64 + ee::kPadW) * sizeof(float);
}
bool canny_tiles_ok(const ee::EdgeArgs& a, bool vec_ok, int th_forced) {
    return fast_eligible(a, vec_ok) && a.W > 128 && a.H % 4 == 0 && a.H >= 16 && th_forced == 0 && g_staging.load() != 3 &&
           a.has_low && a.has_high && a.hyst;
}

// Row-streaming kernels (ee_edge_stream.cuh): full-width bands, hysteresis mode, C == 3.  Measured on B200
// (profiles/r2a_stream_sweep.txt, 3x224x224, us streaming / tiles): backward B = 512: 440 / 644, 128: 132 / 167, 64: 78 / 87,
// 32: 55 / 44; forward 512: 229 / 221, 128: 79 / 64.  So by default they take the backward of batches with >= 64 * 224 image
// rows; staging 8 forces them for every eligible shape and direction (tests), staging 9 disables them.
bool canny_stream_ok(const ee::EdgeArgs& a, bool vec_ok, bool nhwc, bool bwd) {
    const int st = g_staging.load();
    if (st == 9 || st == 1 || st == 3 || st == 4 || nhwc) return false;
    if (!(fast_eligible(a, vec_ok) && a.C == 3 && a.has_low && a.has_high && a.hyst && a.W >= 8 && a.W <= 1024 && a.H >= 2)) return false;
    if (st == 8) return true;
    // a role's warps cover W / 4 column groups: 224 px fills 56 of 64 lanes, 288 px only 72 of 96 (and its 192-thread CTAs fit
    // twice per SM) -- measured 516 us streaming vs 477 us tiles at 256x3x288x288, so such widths stay on the tiles
    const int gx = a.W / 4, lanes = ((gx + 31) / 32) * 32;
    return bwd && a.W > 128 && 8 * gx >= 7 * lanes && (long long)a.B * a.H >= 64LL * 224;
}

// CannyFilter_step125_1 backward through the row-streaming kernel (VAR = 0): OPT-IN (staging 8; parity-tested).  Measured
// on B200 (profiles/r2m_step125_stream.txt, us streaming / chunk-aligned tiles): 512x3x224x224 381 / 325, 128: 125 / 89,
// 64: 70 / 52 -- without the suppression and hysteresis work the pipeline still costs 381 us, i.e. the per-row cost of the
// streaming skeleton itself (one CTA barrier, ring rotations and range guards per row), not the Canny stages, is what bounds
// that kernel family; the halo tiles stay the default for step125.
bool step125_stream_ok(const ee::EdgeArgs& a, bool vec_ok, bool nhwc) {
    if (g_staging.load() != 8 || nhwc) return false;
    return fast_eligible(a, vec_ok) && a.C == 3 && a.has_high && a.W >= 8 && a.W <= 1024 && a.H >= 2 && a.g_x != nullptr;
}

// Tensor map of x as [B*C, H, W] fp32 with a (64 + 8) x (TH + 8) x 1 box for the TMA-staged tile kernel.  The driver
// entry point is looked up at run time (no link-time dependency on libcuda).
int make_x_tensor_map(CUtensorMap* map, const float* x, int B, int C, int H, int W, int box_rows, int plane_w = 64) {
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static std::atomic<encode_fn> cached{nullptr};
    encode_fn fn = cached.load();
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) return fail(EE_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled is not available");
        fn = (encode_fn)p;
        cached.store(fn);
    }
    const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B * (cuuint64_t)C};
    const cuuint64_t strides[2] = {(cuuint64_t)W * sizeof(float), (cuuint64_t)H * W * sizeof(float)};
    const cuuint32_t box[3] = {(cuuint32_t)(plane_w + ee::kPadW), (cuuint32_t)box_rows, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(x), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(EE_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return EE_OK;
}

template <typename K>
int launch_tiles(K kernel, const Launch& L, int B, const ee::FastArgs& a, const CUtensorMap& map, cudaStream_t s, const char* name) {
    if (int rc = ensure_smem(kernel, L.smem)) return rc;
    const long long grid = (long long)B * L.tiles;
    if (grid > 0x7fffffffLL) return fail(EE_ERR_TOO_LARGE, "too many tiles");
    kernel<<<(unsigned)grid, L.threads, L.smem, s>>>(a, map);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, name);
    return EE_OK;
}

bool fast_eligible(const ee::EdgeArgs& a, bool vec_ok) {
    if (g_staging.load() == 1 || !vec_ok || a.W < 8 || a.H < 4) return false;
    if (a.nan_compat && a.g_in != nullptr) return false;      // reference-NaN backward: shape-generic kernels only
    if (a.strided) return false;                              // non-dense strides: shape-generic kernels only
    return std::isfinite(a.high) && std::isfinite(a.alpha) && std::isfinite(a.low) && fabsf(a.high) < 1e18f &&
           fabsf(a.alpha) < 1e18f;
}

// Tiny-ImageNet configuration of the full Canny / BPDA filter: 3x64x64, both thresholds, hysteresis on --
// compiled with variant and output mode as constants
bool hot_canny(const ee::FastArgs& f, const Launch& L) {
    return L.tiles_x == 1 && f.e.W == 64 && f.e.C == 3 && f.e.has_low && f.e.has_high && f.e.hyst;
}

void fill_fast(ee::FastArgs& f, const ee::EdgeArgs& a, const Launch& L) {
    memset(&f.x_map, 0, sizeof(f.x_map));
    f.e = a;
    f.e.TH = L.TH; f.e.tiles_per_img = L.tiles; f.e.GX = L.GX; f.e.RY = L.RY;
    f.hi_cut = cut_gt(a.high);
    f.w_cut = cut_gt(1.001f);
    f.a_cut = cut_lt(a.alpha);
    // u > hi_cut and u >= a_cut  <=>  u > max(hi_cut, nextbelow(a_cut))
    f.e_cut = f.hi_cut;
    if (f.a_cut > 0.0f) {
        const float below = nextafterf(f.a_cut, -INFINITY);
        if (below > f.e_cut) f.e_cut = below;
    }
    f.zero_val = (0.0f > a.high) ? 1.0f : 0.0f;
    f.Wp = L.planeW + ee::kPadW;
    f.TW = L.TW;
    f.tiles_x = L.tiles_x;
    f.halo = L.halo;
}

// Specialised instantiations <NC, BLEND, rows/thread, plane width WT, global width WG> for the hot shapes
// (Tiny-ImageNet 64, CIFAR 32, MNIST 28 as single column tiles; 64-wide planes of column-tiled 224 px
// images); everything else runs the runtime-width instantiation <.., 0, 0> of the same kernel.
#define EE_DISPATCH_FAST_NHWC(KERNEL, L, B, f, s, name)                                           \
    do {                                                                                          \
        if (whole_image(L, f, 64)) return launch_fast_even(KERNEL<3, true, 4, 64, 64, true, 64>, L, B, f, s, name); \
        if ((L).tiles_x == 1 && (f).e.W == 64) return launch_fast(KERNEL<3, true, 4, 64, 64, true>, L, B, f, s, name); \
        return launch_fast(KERNEL<3, true, 4, 0, 0, true>, L, B, f, s, name);                     \
    } while (0)

#define EE_DISPATCH_FAST(KERNEL, BLEND, L, B, f, s, name)                                         \
    do {                                                                                          \
        const int W_ = (f).e.W, C_ = (f).e.C;                                                     \
        if ((L).tiles_x == 1) {                                                                   \
            if (C_ == 3 && whole_image(L, f, 64)) return launch_fast_even(KERNEL<3, BLEND, 4, 64, 64, false, 64>, L, B, f, s, name); \
            if (C_ == 3 && whole_image(L, f, 32)) return launch_fast_even(KERNEL<3, BLEND, 4, 32, 32, false, 32>, L, B, f, s, name); \
            if (C_ == 1 && whole_image(L, f, 28)) return launch_fast_even(KERNEL<1, BLEND, 4, 28, 28, false, 28>, L, B, f, s, name); \
            if (C_ == 3 && W_ == 64) return launch_fast(KERNEL<3, BLEND, 4, 64, 64>, L, B, f, s, name); \
            if (C_ == 3 && W_ == 32) return launch_fast(KERNEL<3, BLEND, 4, 32, 32>, L, B, f, s, name); \
            if (C_ == 1 && W_ == 28) return launch_fast(KERNEL<1, BLEND, 4, 28, 28>, L, B, f, s, name); \
            if (C_ == 1 && W_ == 32) return launch_fast(KERNEL<1, BLEND, 4, 32, 32>, L, B, f, s, name); \
        } else if (C_ == 3 && (L).planeW == 64) {                                                 \
            return launch_fast(KERNEL<3, BLEND, 4, 64, 0>, L, B, f, s, name);                     \
        }                                                                                         \
        if (C_ == 3) return launch_fast(KERNEL<3, BLEND, 4, 0, 0>, L, B, f, s, name);             \
        if (C_ == 1) return launch_fast(KERNEL<1, BLEND, 4, 0, 0>, L, B, f, s, name);             \
        return launch_fast(KERNEL<0, BLEND, 4, 0, 0>, L, B, f, s, name);                          \
    } while (0)

// rows_fixed / rows_per_th: the kernel needs (rows_per_th*TH + rows_fixed) plane rows of W floats.
int plan(int H, int W, bool vec_ok, int rows_per_th, int rows_fixed, int max_halo_rows, int budget_bytes,
         int forced_th, Launch& L) {
    L.vec = vec_ok ? 4 : 1;
    const int G = (W + L.vec - 1) / L.vec;
    const size_t row_bytes = (size_t)W * sizeof(float);
    auto smem_of = [&](int th) { return (size_t)(rows_per_th * th + rows_fixed) * row_bytes; };
    int th;
    if (forced_th > 0) {
        th = forced_th < H ? forced_th : H;
    } else {
        long fit = ((long)(budget_bytes / row_bytes) - rows_fixed) / rows_per_th;
        th = (int)(fit < 1 ? 1 : (fit > H ? H : fit));
        const int tiles = (H + th - 1) / th;
        th = (H + tiles - 1) / tiles;   // equalise the strips
    }
    while (th > 1 && smem_of(th) > (size_t)kMaxSmem) --th;
    if (smem_of(th) > (size_t)kMaxSmem)
        return fail(EE_ERR_TOO_LARGE, "one row strip of width %d needs %zu B of shared memory (> %d)", W, smem_of(th), kMaxSmem);
    L.TH = th;
    L.tiles = (H + th - 1) / th;
    L.smem = smem_of(th);
    if (G >= kThreads) { L.GX = kThreads; L.RY = 1; }
    else {
        L.GX = G;
        int ry = kThreads / G;
        const int max_rows = th + max_halo_rows;
        if (ry > max_rows) ry = max_rows;
        if (ry < 1) ry = 1;
        L.RY = ry;
    }
    L.threads = ((L.GX * L.RY + 31) / 32) * 32;
    return EE_OK;
}

template <typename K>
int launch(K kernel, const Launch& L, int B, const ee::EdgeArgs& a, cudaStream_t s, const char* name) {
    if (int rc = ensure_smem(kernel, L.smem)) return rc;
    const long long grid = (long long)B * L.tiles;
    if (grid > 0x7fffffffLL) return fail(EE_ERR_TOO_LARGE, "too many tiles");
    kernel<<<(unsigned)grid, L.threads, L.smem, s>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, name);
    return EE_OK;
}

int fill_args(ee::EdgeArgs& a, int B, int C, int H, int W, const EEParams* p, float w) {
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return fail(EE_ERR_INVALID_ARG, "non-positive shape %dx%dx%dx%d", B, C, H, W);
    float c0, c1, c2;
    int rc = check_params(p, c0, c1, c2);
    if (rc) return rc;
    memset(&a, 0, sizeof(a));
    a.B = B; a.C = C; a.H = H; a.W = W;
    a.c0 = c0; a.c1 = c1; a.c2 = c2; a.fC = (float)C;
    a.alpha = p->alpha; a.low = p->low_thr; a.high = p->high_thr; a.w = w;
    a.variant = p->variant; a.has_low = p->has_low != 0; a.has_high = p->has_high != 0;
    // CannyFilter_step125_1 ignores low_threshold / hysteresis (core.py:549-585): normalise them here so that every entry
    // point and every predicate below sees the same request whatever the caller left in those fields
    if (p->variant == EE_VARIANT_STEP125) a.has_low = 0;
    a.hyst = (p->variant != EE_VARIANT_STEP125) && p->hysteresis != 0;
    a.nan_compat = (p->flags & EE_FLAG_NAN_COMPAT) != 0;
    const int64_t hw = (int64_t)H * W;
    const ee::EEStride3 dense = {(int64_t)C * hw, hw, (int64_t)W}, dense1 = {hw, hw, (int64_t)W};
    a.sx = a.sbase = a.sg = a.sout = a.sgx = a.sgbase = dense;
    a.sedge = dense1;
    a.strided = 0;
    return EE_OK;
}

#define EE_DISPATCH(KERNEL, BLEND, L, B, a, s, name)                                              \
    do {                                                                                          \
        if ((L).vec == 4) {                                                                       \
            if ((a).C == 3) return launch(KERNEL<4, 3, BLEND>, L, B, a, s, name);                 \
            if ((a).C == 1) return launch(KERNEL<4, 1, BLEND>, L, B, a, s, name);                 \
            return launch(KERNEL<4, 0, BLEND>, L, B, a, s, name);                                 \
        }                                                                                         \
        if ((a).C == 3) return launch(KERNEL<1, 3, BLEND>, L, B, a, s, name);                     \
        if ((a).C == 1) return launch(KERNEL<1, 1, BLEND>, L, B, a, s, name);                     \
        return launch(KERNEL<1, 0, BLEND>, L, B, a, s, name);                                     \
    } while (0)

#define EE_NHWC_MSG "NHWC needs the fused blend entry point, C == 3, W %% 4 == 0, 16-byte aligned tensors"

}  // namespace

// =================================================================================================
// kernel-family dispatchers
// =================================================================================================
#if EE_HAS(2)
int ee_shared::fwd_step125(ee::EdgeArgs& a, const EEParams* p, int B, bool blend, bool vec_ok, bool nhwc, cudaStream_t s) {
    const int C = a.C, H = a.H, W = a.W;
    Launch L;
    int rc;
    if (nhwc) {
        // channels_last is implemented by the tuned fused kernels for C = 3; anything else must be
        // converted by the caller (the Python wrapper does), never silently mis-read
        if (!(blend && C == 3 && fast_eligible(a, vec_ok))) return fail(EE_ERR_UNSUPPORTED, EE_NHWC_MSG);
        rc = plan_fast(H, W, 4, 2, 6, 4, 40 * 1024, g_th_fwd.load(), L, 4, 512);
        if (rc) return rc;
        ee::FastArgs f;
        fill_fast(f, a, L);
        EE_DISPATCH_FAST_NHWC(ee::edge_fwd_step125_fast, L, B, f, s, "edge_fwd_step125_fast_nhwc");
    }
    if (fast_eligible(a, vec_ok)) {
        rc = plan_fast(H, W, 4, 2, 6, 4, 40 * 1024, g_th_fwd.load(), L, 4, 512);
        if (rc) return rc;
        ee::FastArgs f;
        fill_fast(f, a, L);
        if (blend) EE_DISPATCH_FAST(ee::edge_fwd_step125_fast, true, L, B, f, s, "edge_fwd_step125_fast");
        else EE_DISPATCH_FAST(ee::edge_fwd_step125_fast, false, L, B, f, s, "edge_fwd_step125_fast");
    }
    rc = plan(H, W, vec_ok, 2, 6, 4, 44 * 1024, g_th_fwd.load(), L);
    if (rc) return rc;
    a.TH = L.TH; a.tiles_per_img = L.tiles; a.GX = L.GX; a.RY = L.RY;
    if (blend) EE_DISPATCH(ee::edge_fwd_step125_kernel, true, L, B, a, s, "edge_fwd_step125");
    else EE_DISPATCH(ee::edge_fwd_step125_kernel, false, L, B, a, s, "edge_fwd_step125");
}
#endif

#if EE_HAS(4)
int ee_shared::fwd_canny(ee::EdgeArgs& a, const EEParams* p, int B, bool blend, bool vec_ok, bool nhwc, cudaStream_t s) {
    const int C = a.C, H = a.H, W = a.W;
    Launch L;
    int rc;
    if (nhwc) {
        if (!(blend && C == 3 && fast_eligible(a, vec_ok))) return fail(EE_ERR_UNSUPPORTED, EE_NHWC_MSG);
        rc = plan_fast(H, W, 4, ee::kCannyFastFwdRowsPerTH, ee::kCannyFastFwdRowsFixed, 8, 62 * 1024, g_th_fwd.load(), L);
        if (rc) return rc;
        ee::FastArgs f;
        fill_fast(f, a, L);
        EE_DISPATCH_FAST_NHWC(ee::edge_fwd_canny_fast, L, B, f, s, "edge_fwd_canny_fast_nhwc");
    }
    if (canny_stream_ok(a, vec_ok, nhwc, false)) return ee_shared::canny_stream(a, p, B, blend, false, s);
    if (canny_tiles_ok(a, vec_ok, g_th_fwd.load())) {
        // wide images, hysteresis mode: chunk-aligned 56 x 56 tiles (ee_edge_canny_tiles.cuh)
        plan_tiles(H, W, L, 56, 3);
        ee::FastArgs f;
        fill_fast(f, a, L);
        const bool cn = (p->variant == EE_VARIANT_CANNY);
        if (C == 3 && blend) return cn ? launch_fast(ee::edge_fwd_canny_tiles<3, true, 4, 64, 1>, L, B, f, s, "edge_fwd_canny_tiles")
                                       : launch_fast(ee::edge_fwd_canny_tiles<3, true, 4, 64, 2>, L, B, f, s, "edge_fwd_canny_tiles");
        if (blend) return launch_fast(ee::edge_fwd_canny_tiles<0, true, 4, 64, 0>, L, B, f, s, "edge_fwd_canny_tiles");
        return launch_fast(ee::edge_fwd_canny_tiles<0, false, 4, 64, 0>, L, B, f, s, "edge_fwd_canny_tiles");
    }
    if (fast_eligible(a, vec_ok)) {
        rc = plan_fast(H, W, 4, ee::kCannyFastFwdRowsPerTH, ee::kCannyFastFwdRowsFixed, 8, 62 * 1024, g_th_fwd.load(), L);
        if (rc) return rc;
        ee::FastArgs f;
        fill_fast(f, a, L);
        if (blend && hot_canny(f, L)) {
            if (whole_image(L, f, 64)) {
                if (p->variant == EE_VARIANT_CANNY)
                    return launch_fast_even(ee::edge_fwd_canny_fast<3, true, 4, 64, 64, false, 64, 1, ee::MODE_HYST>, L, B, f, s, "edge_fwd_canny_fast");
                return launch_fast_even(ee::edge_fwd_canny_fast<3, true, 4, 64, 64, false, 64, 2, ee::MODE_HYST>, L, B, f, s, "edge_fwd_canny_fast");
            }
            if (p->variant == EE_VARIANT_CANNY)
                return launch_fast(ee::edge_fwd_canny_fast<3, true, 4, 64, 64, false, 0, 1, ee::MODE_HYST>, L, B, f, s, "edge_fwd_canny_fast");
            return launch_fast(ee::edge_fwd_canny_fast<3, true, 4, 64, 64, false, 0, 2, ee::MODE_HYST>, L, B, f, s, "edge_fwd_canny_fast");
        }
        if (blend) EE_DISPATCH_FAST(ee::edge_fwd_canny_fast, true, L, B, f, s, "edge_fwd_canny_fast");
        else EE_DISPATCH_FAST(ee::edge_fwd_canny_fast, false, L, B, f, s, "edge_fwd_canny_fast");
    }
    rc = plan(H, W, vec_ok, ee::kCannyFwdRowsPerTH, ee::kCannyFwdRowsFixed, 8, 64 * 1024, g_th_fwd.load(), L);
    if (rc) return rc;
    a.TH = L.TH; a.tiles_per_img = L.tiles; a.GX = L.GX; a.RY = L.RY;
    if (blend) EE_DISPATCH(ee::edge_fwd_canny_kernel, true, L, B, a, s, "edge_fwd_canny");
    else EE_DISPATCH(ee::edge_fwd_canny_kernel, false, L, B, a, s, "edge_fwd_canny");
}
#endif

#if EE_HAS(3)
int ee_shared::bwd_step125(ee::EdgeArgs& a, const EEParams* p, int B, bool blend, bool vec_ok, bool nhwc, cudaStream_t s) {
    const int C = a.C, H = a.H, W = a.W;
    const float* x = a.x;
    Launch L;
    int rc;
    if (nhwc) {
        if (!(blend && C == 3 && fast_eligible(a, vec_ok))) return fail(EE_ERR_UNSUPPORTED, EE_NHWC_MSG);
        rc = plan_fast(H, W, 4, 3, 18, 8, 62 * 1024, g_th_bwd.load(), L);
        if (rc) return rc;
        ee::FastArgs f;
        fill_fast(f, a, L);
        EE_DISPATCH_FAST_NHWC(ee::edge_bwd_step125_fast, L, B, f, s, "edge_bwd_step125_fast_nhwc");
    }
    if (fast_eligible(a, vec_ok) && g_staging.load() == 5 && C == 3 && H == 224 && W == 224) {
        // opt-in (staging 5; measured equal to the strip kernels, ee_edge_cluster.cuh): ImageNet size as one cluster of
        // 8 CTAs x 28 rows per image, halo rows exchanged through distributed shared memory
        if (blend) return launch_cluster(ee::edge_bwd_step125_cluster<3, true, 4, 224, 28, 8>, 224, 28, 8, B, a, s, "edge_bwd_step125_cluster");
        return launch_cluster(ee::edge_bwd_step125_cluster<3, false, 4, 224, 28, 8>, 224, 28, 8, B, a, s, "edge_bwd_step125_cluster");
    }
    if (step125_stream_ok(a, vec_ok, nhwc)) return ee_shared::canny_stream(a, p, B, blend, true, s);
    if (fast_eligible(a, vec_ok) && W > 128 && H % 4 == 0 && H >= 16 && g_th_bwd.load() == 0 && g_staging.load() != 3) {
        // wide images: chunk-aligned 56 x 56 tiles, one chunk per thread, no row guards (staging 3 = the older strip path;
        // staging 6 = x tiles staged by TMA tensor copies instead of LDGs, C == 3)
        plan_tiles(H, W, L);
        ee::FastArgs f;
        fill_fast(f, a, L);
        CUtensorMap map;
        memset(&map, 0, sizeof(map));
        const bool full_tiles = (L.TH == 56 && L.TW == 56);     // measured: TMA staging +3 % at 224 px, -2.5 % at 288 px (48 x 48 tiles)
        if (C == 3 && g_staging.load() != 7 && (g_staging.load() == 6 || full_tiles) &&
            make_x_tensor_map(&map, x, B, C, H, W, L.TH + 8) == EE_OK) {
            if (blend) return launch_tiles(ee::edge_bwd_step125_tiles<3, true, 4, 64, true>, L, B, f, map, s, "edge_bwd_step125_tiles_tma");
            return launch_tiles(ee::edge_bwd_step125_tiles<3, false, 4, 64, true>, L, B, f, map, s, "edge_bwd_step125_tiles_tma");
        }
        if (C == 3) {
            if (blend) return launch_tiles(ee::edge_bwd_step125_tiles<3, true, 4, 64>, L, B, f, map, s, "edge_bwd_step125_tiles");
            return launch_tiles(ee::edge_bwd_step125_tiles<3, false, 4, 64>, L, B, f, map, s, "edge_bwd_step125_tiles");
        }
        if (blend) return launch_tiles(ee::edge_bwd_step125_tiles<0, true, 4, 64>, L, B, f, map, s, "edge_bwd_step125_tiles");
        return launch_tiles(ee::edge_bwd_step125_tiles<0, false, 4, 64>, L, B, f, map, s, "edge_bwd_step125_tiles");
    }
    if (fast_eligible(a, vec_ok)) {
        rc = plan_fast(H, W, 4, 3, 18, 8, 62 * 1024, g_th_bwd.load(), L);
        if (rc) return rc;
        ee::FastArgs f;
        fill_fast(f, a, L);
        if (C == 3 && whole_image(L, f, 64) && g_staging.load() == 6 && make_x_tensor_map(&f.x_map, x, B, C, H, W, 64, 64) == EE_OK) {
            // opt-in (staging 6): the three channel planes of x staged by TMA tensor copies.  Measured at 4096x3x64x64:
            // 5.66 TB/s vs 6.07 TB/s with LDGs summed in registers (every thread waits for all three planes, and the sum
            // costs three LDS.128 per row), so the LDG path stays the default for whole-image tiles.
            if (blend) return launch_fast_even(ee::edge_bwd_step125_fast<3, true, 4, 64, 64, false, 64, true>, L, B, f, s, "edge_bwd_step125_fast_tma");
            return launch_fast_even(ee::edge_bwd_step125_fast<3, false, 4, 64, 64, false, 64, true>, L, B, f, s, "edge_bwd_step125_fast_tma");
        }
        if (blend) EE_DISPATCH_FAST(ee::edge_bwd_step125_fast, true, L, B, f, s, "edge_bwd_step125_fast");
        else EE_DISPATCH_FAST(ee::edge_bwd_step125_fast, false, L, B, f, s, "edge_bwd_step125_fast");
    }
    rc = plan(H, W, vec_ok, 3, 18, 8, 56 * 1024, g_th_bwd.load(), L);
    if (rc) return rc;
    a.TH = L.TH; a.tiles_per_img = L.tiles; a.GX = L.GX; a.RY = L.RY;
    if (blend) EE_DISPATCH(ee::edge_bwd_step125_kernel, true, L, B, a, s, "edge_bwd_step125");
    else EE_DISPATCH(ee::edge_bwd_step125_kernel, false, L, B, a, s, "edge_bwd_step125");
}
#endif

#if EE_HAS(5)
int ee_shared::bwd_canny(ee::EdgeArgs& a, const EEParams* p, int B, bool blend, bool vec_ok, bool nhwc, cudaStream_t s) {
    const int C = a.C, H = a.H, W = a.W;
    Launch L;
    int rc;
    if (nhwc) {
        if (!(blend && C == 3 && fast_eligible(a, vec_ok))) return fail(EE_ERR_UNSUPPORTED, EE_NHWC_MSG);
        rc = plan_fast(H, W, 4, ee::kCannyFastBwdRowsPerTH, ee::kCannyFastBwdRowsFixed, 12, 104 * 1024, g_th_bwd.load(), L, 8, 128, 48);
        if (rc) return rc;
        L.even_planes = 4;      // whole-image Canny backward: gy1 reuses the blurred plane's region
        ee::FastArgs f;
        fill_fast(f, a, L);
        EE_DISPATCH_FAST_NHWC(ee::edge_bwd_canny_fast, L, B, f, s, "edge_bwd_canny_fast_nhwc");
    }
    if (a.g_x != nullptr && canny_stream_ok(a, vec_ok, nhwc, true)) return ee_shared::canny_stream(a, p, B, blend, true, s);
    if (canny_tiles_ok(a, vec_ok, g_th_bwd.load())) {
        // wide images, hysteresis mode: chunk-aligned 48 x 48 tiles with an 8-pixel halo (ee_edge_canny_tiles.cuh)
        plan_tiles(H, W, L, 48, 4);
        ee::FastArgs f;
        fill_fast(f, a, L);
        const bool cn = (p->variant == EE_VARIANT_CANNY);
        if (C == 3 && blend) return cn ? launch_fast(ee::edge_bwd_canny_tiles<3, true, 4, 64, 1>, L, B, f, s, "edge_bwd_canny_tiles")
                                       : launch_fast(ee::edge_bwd_canny_tiles<3, true, 4, 64, 2>, L, B, f, s, "edge_bwd_canny_tiles");
        if (blend) return launch_fast(ee::edge_bwd_canny_tiles<0, true, 4, 64, 0>, L, B, f, s, "edge_bwd_canny_tiles");
        return launch_fast(ee::edge_bwd_canny_tiles<0, false, 4, 64, 0>, L, B, f, s, "edge_bwd_canny_tiles");
    }
    if (fast_eligible(a, vec_ok)) {
        rc = plan_fast(H, W, 4, ee::kCannyFastBwdRowsPerTH, ee::kCannyFastBwdRowsFixed, 12, 104 * 1024, g_th_bwd.load(), L, 8, 128, 48);
        if (rc) return rc;
        L.even_planes = 4;               // whole-image Canny backward: gy1 reuses the blurred plane's region
        ee::FastArgs f;
        fill_fast(f, a, L);
        if (blend && hot_canny(f, L)) {
            if (whole_image(L, f, 64)) {
                if (p->variant == EE_VARIANT_CANNY)
                    return launch_fast_even(ee::edge_bwd_canny_fast<3, true, 4, 64, 64, false, 64, 1, ee::MODE_HYST>, L, B, f, s, "edge_bwd_canny_fast");
                return launch_fast_even(ee::edge_bwd_canny_fast<3, true, 4, 64, 64, false, 64, 2, ee::MODE_HYST>, L, B, f, s, "edge_bwd_canny_fast");
            }
            if (p->variant == EE_VARIANT_CANNY)
                return launch_fast(ee::edge_bwd_canny_fast<3, true, 4, 64, 64, false, 0, 1, ee::MODE_HYST>, L, B, f, s, "edge_bwd_canny_fast");
            return launch_fast(ee::edge_bwd_canny_fast<3, true, 4, 64, 64, false, 0, 2, ee::MODE_HYST>, L, B, f, s, "edge_bwd_canny_fast");
        }
        if (blend) EE_DISPATCH_FAST(ee::edge_bwd_canny_fast, true, L, B, f, s, "edge_bwd_canny_fast");
        else EE_DISPATCH_FAST(ee::edge_bwd_canny_fast, false, L, B, f, s, "edge_bwd_canny_fast");
    }
    rc = plan(H, W, vec_ok, ee::kCannyBwdRowsPerTH, ee::kCannyBwdRowsFixed, 12, 80 * 1024, g_th_bwd.load(), L);
    if (rc) return rc;
    a.TH = L.TH; a.tiles_per_img = L.tiles; a.GX = L.GX; a.RY = L.RY;
    if (blend) EE_DISPATCH(ee::edge_bwd_canny_kernel, true, L, B, a, s, "edge_bwd_canny");
    else EE_DISPATCH(ee::edge_bwd_canny_kernel, false, L, B, a, s, "edge_bwd_canny");
}
#endif

#if EE_HAS(6)
namespace {
template <typename K>
int launch_stream(K kernel, const ee::StreamArgs& sa, unsigned grid, int threads, size_t smem, cudaStream_t s, const char* name) {
    if (int rc = ensure_smem(kernel, smem)) return rc;
    kernel<<<grid, threads, smem, s>>>(sa);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, name);
    return EE_OK;
}
template <bool BLEND, int VAR, bool BWD>
int stream_by_width(const ee::StreamArgs& sa, int W, unsigned grid, int threads, size_t smem, cudaStream_t s) {
    const char* name = BWD ? "edge_bwd_canny_stream" : "edge_fwd_canny_stream";
    // ImageNet sizes with the width as a compile-time constant; <.., WT, max threads, min CTAs per SM>
    if (W == 224) return launch_stream(ee::edge_canny_stream<3, BLEND, VAR, BWD, 224, 128, BWD ? 4 : 6>, sa, grid, threads, smem, s, name);
    if (W == 288) return launch_stream(ee::edge_canny_stream<3, BLEND, VAR, BWD, 288, 192, BWD ? 2 : 4>, sa, grid, threads, smem, s, name);
    if (threads <= 64) return launch_stream(ee::edge_canny_stream<3, BLEND, VAR, BWD, 0, 64, 8>, sa, grid, threads, smem, s, name);
    if (threads <= 128) return launch_stream(ee::edge_canny_stream<3, BLEND, VAR, BWD, 0, 128, 4>, sa, grid, threads, smem, s, name);
    if (threads <= 256) return launch_stream(ee::edge_canny_stream<3, BLEND, VAR, BWD, 0, 256, 2>, sa, grid, threads, smem, s, name);
    return launch_stream(ee::edge_canny_stream<3, BLEND, VAR, BWD, 0, 512, 1>, sa, grid, threads, smem, s, name);
}
}  // namespace

int ee_shared::canny_stream(ee::EdgeArgs& a, const EEParams* p, int B, bool blend, bool bwd, cudaStream_t s) {
    const int H = a.H, W = a.W, GX = W / 4;
    ee::StreamArgs sa;
    Launch L;
    L.vec = 4; L.TH = H; L.tiles = 1; L.GX = GX; L.RY = 1; L.TW = W; L.tiles_x = 1; L.planeW = W; L.halo = 0;
    fill_fast(sa.f, a, L);
    const int threads = 2 * (((GX + 31) / 32) * 32);          // front + back role
    const int rows = 3 + (blend ? 3 : 0) + (bwd ? (blend ? 3 : 1) : 0);
    const size_t smem = ee::stream_smem_bytes(W, rows, bwd);
    if (smem > (size_t)kMaxSmem) return fail(EE_ERR_TOO_LARGE, "row-streaming kernel: a %d-column row ring does not fit in shared memory", W);
    // band height: about two bands per resident-CTA slot (4 CTAs per SM), but never shorter than 28 rows -- a band warms
    // its pipeline up on 2 * halo = 8 / 12 extra rows; measured flat between 56 and 224 rows at B = 512
    int bh = bwd ? g_th_bwd.load() : g_th_fwd.load();
    if (bh <= 0) {
        const int slots = 148 * 4;
        int bands = (2 * slots + B - 1) / B;
        if (bands < 1) bands = 1;
        bh = (H + bands - 1) / bands;
        if (bh < 28) bh = 28;
    }
    if (bh > H) bh = H;
    sa.BH = bh;
    sa.bands_per_img = (H + bh - 1) / bh;
    const long long grid = (long long)B * sa.bands_per_img;
    if (grid > 0x7fffffffLL) return fail(EE_ERR_TOO_LARGE, "too many bands");
    const bool cn = (p->variant == EE_VARIANT_CANNY);
    if (p->variant == EE_VARIANT_STEP125) {          // backward only (the forward strips are at the HBM roofline already)
        if (!bwd) return fail(EE_ERR_UNSUPPORTED, "row-streaming kernel: CannyFilter_step125_1 is wired for the backward only");
        return blend ? stream_by_width<true, 0, true>(sa, W, (unsigned)grid, threads, smem, s) : stream_by_width<false, 0, true>(sa, W, (unsigned)grid, threads, smem, s);
    }
    if (bwd) {
        if (blend) return cn ? stream_by_width<true, 1, true>(sa, W, (unsigned)grid, threads, smem, s) : stream_by_width<true, 2, true>(sa, W, (unsigned)grid, threads, smem, s);
        return cn ? stream_by_width<false, 1, true>(sa, W, (unsigned)grid, threads, smem, s) : stream_by_width<false, 2, true>(sa, W, (unsigned)grid, threads, smem, s);
    }
    if (blend) return cn ? stream_by_width<true, 1, false>(sa, W, (unsigned)grid, threads, smem, s) : stream_by_width<true, 2, false>(sa, W, (unsigned)grid, threads, smem, s);
    return cn ? stream_by_width<false, 1, false>(sa, W, (unsigned)grid, threads, smem, s) : stream_by_width<false, 2, false>(sa, W, (unsigned)grid, threads, smem, s);
}
#endif

#if EE_HAS(1)
namespace {

// How a caller-described tensor can be read: dense in a layout the tuned kernels know, rows contiguous (generic kernels
// on the strides), or not at all.
enum StrideKind { SK_DENSE_NCHW, SK_DENSE_NHWC, SK_ROWS, SK_BAD };
StrideKind classify(const EEStrides* s, int B, int C, int H, int W, int layout, ee::EEStride3& out) {
    const int64_t hw = (int64_t)H * W;
    if (!s) {
        out = {(int64_t)C * hw, hw, (int64_t)W};
        return (layout == EE_LAYOUT_NHWC && C > 1) ? SK_DENSE_NHWC : SK_DENSE_NCHW;
    }
    // strides of size-1 dimensions carry no information (torch leaves them arbitrary)
    const int64_t sn = (B == 1) ? (int64_t)C * hw : s->n, sc = (C == 1) ? hw : s->c, sh = (H == 1) ? W : s->h, sw = (W == 1) ? 1 : s->w;
    out = {sn, sc, sh};
    if (sw == 1 && sh == W && sc == hw && sn == (int64_t)C * hw) return SK_DENSE_NCHW;
    if (C > 1 && sc == 1 && sw == C && sh == (int64_t)W * C && (B == 1 || s->n == hw * C)) return SK_DENSE_NHWC;
    if (sw == 1) return SK_ROWS;
    return SK_BAD;
}
bool stride_vec_ok(const ee::EEStride3& s) { return ((s.b | s.c | s.h) & 3) == 0; }

struct TensorDesc { const void* ptr; const EEStrides* s; int C; ee::EEStride3* dst; const char* name; };

// Fills the strides of every described tensor; returns the common kind (dense NCHW / dense NHWC / rows), or fails.
int resolve_strides(TensorDesc* t, int n, int B, int H, int W, int layout, StrideKind& kind, bool& vec_ok) {
    kind = SK_DENSE_NCHW;
    bool any_rows = false, any_nhwc = false, any_nchw = false;
    for (int i = 0; i < n; ++i) {
        if (!t[i].ptr) continue;
        const StrideKind k = classify(t[i].s, B, t[i].C, H, W, layout, *t[i].dst);
        if (k == SK_BAD) return fail(EE_ERR_UNSUPPORTED, "%s: column stride must be 1 (or the tensor dense channels_last); make a contiguous copy", t[i].name);
        if (k == SK_ROWS) any_rows = true;
        else if (k == SK_DENSE_NHWC) any_nhwc = true;
        else if (t[i].C > 1) any_nchw = true;
        vec_ok = vec_ok && stride_vec_ok(*t[i].dst);
    }
    if (any_nhwc && (any_rows || any_nchw)) return fail(EE_ERR_UNSUPPORTED, "mixed memory layouts: every image tensor must be channels_last, or none");
    kind = any_rows ? SK_ROWS : (any_nhwc ? SK_DENSE_NHWC : SK_DENSE_NCHW);
    return EE_OK;
}

int edge_forward(const float* x, const float* base, float* out, float* edge, int B, int C, int H, int W,
                 const EEParams* p, float w, bool blend, void* stream, const EEStrides* xs = nullptr,
                 const EEStrides* bs = nullptr, const EEStrides* os = nullptr, const EEStrides* es = nullptr, bool strided_api = false) {
    ee::EdgeArgs a;
    int rc = fill_args(a, B, C, H, W, p, w);
    if (rc) return rc;
    if (!x) return fail(EE_ERR_INVALID_ARG, "x is null");
    if (blend && (!base || !out)) return fail(EE_ERR_INVALID_ARG, "base/out is null");
    if (!blend && !edge) return fail(EE_ERR_INVALID_ARG, "edge is null");
    a.x = x; a.base = base; a.out = out; a.edge = edge;
    bool vec_ok = (W % 4 == 0) && aligned16(x) && aligned16(base) && aligned16(out) && aligned16(edge);
    bool nhwc = (p->layout == EE_LAYOUT_NHWC) && C > 1;       // with one channel the layouts coincide
    if (strided_api) {
        TensorDesc t[4] = {{x, xs, C, &a.sx, "x"}, {base, bs, C, &a.sbase, "base"}, {out, os, C, &a.sout, "out"}, {edge, es, 1, &a.sedge, "edge"}};
        StrideKind kind;
        rc = resolve_strides(t, 4, B, H, W, p->layout, kind, vec_ok);
        if (rc) return rc;
        nhwc = (kind == SK_DENSE_NHWC);
        a.strided = (kind == SK_ROWS);
        if (nhwc && a.edge && classify(es, B, 1, H, W, EE_LAYOUT_NCHW, a.sedge) != SK_DENSE_NCHW)
            return fail(EE_ERR_UNSUPPORTED, "edge must be dense next to channels_last images");
    }
    if (p->variant == EE_VARIANT_STEP125) return ee_shared::fwd_step125(a, p, B, blend, vec_ok, nhwc, (cudaStream_t)stream);
    return ee_shared::fwd_canny(a, p, B, blend, vec_ok, nhwc, (cudaStream_t)stream);
}

int edge_backward(const float* g_in, const float* x, const float* base, float* g_x, float* g_base, int B, int C,
                  int H, int W, const EEParams* p, float w, bool blend, void* stream, const EEStrides* gs = nullptr,
                  const EEStrides* xs = nullptr, const EEStrides* bs = nullptr, const EEStrides* gxs = nullptr,
                  const EEStrides* gbs = nullptr, bool strided_api = false) {
    ee::EdgeArgs a;
    int rc = fill_args(a, B, C, H, W, p, w);
    if (rc) return rc;
    if (!x || !g_in) return fail(EE_ERR_INVALID_ARG, "x/g is null");
    if (blend && !base) return fail(EE_ERR_INVALID_ARG, "base is null");
    if (!blend && !g_x) return fail(EE_ERR_INVALID_ARG, "g_x is null");
    if (blend && !g_x && !g_base) return EE_OK;   // nothing requested
    a.x = x; a.base = base; a.g_in = g_in; a.g_x = g_x; a.g_base = g_base;
    if (!blend) a.sg = a.sedge;                      // module-level backward: the upstream gradient is the [B,1,H,W] edge gradient
    bool vec_ok = (W % 4 == 0) && aligned16(x) && aligned16(base) && aligned16(g_in) && aligned16(g_x) && aligned16(g_base);
    bool nhwc = (p->layout == EE_LAYOUT_NHWC) && C > 1;
    if (strided_api) {
        TensorDesc t[5] = {{g_in, gs, blend ? C : 1, &a.sg, "grad"}, {x, xs, C, &a.sx, "x"}, {base, bs, C, &a.sbase, "base"},
                           {g_x, gxs, C, &a.sgx, "g_x"}, {g_base, gbs, C, &a.sgbase, "g_base"}};
        StrideKind kind;
        rc = resolve_strides(t, 5, B, H, W, p->layout, kind, vec_ok);
        if (rc) return rc;
        nhwc = (kind == SK_DENSE_NHWC);
        a.strided = (kind == SK_ROWS);
    }
    if (p->variant == EE_VARIANT_STEP125) return ee_shared::bwd_step125(a, p, B, blend, vec_ok, nhwc, (cudaStream_t)stream);
    return ee_shared::bwd_canny(a, p, B, blend, vec_ok, nhwc, (cudaStream_t)stream);
}

// ---- elementwise launcher ----------------------------------------------------------------
template <int NIN, typename F>
int launch_ew(const float* i0, const float* i1, const float* i2, const float* i3, const float* i4, float* out,
              int64_t n, F f, void* stream, const char* name) {
    if (n < 0) return fail(EE_ERR_INVALID_ARG, "negative element count");
    if (n == 0) return EE_OK;
    const float* in[5] = {i0, i1, i2, i3, i4};
    int vec_ok = aligned16(out);
    for (int q = 0; q < NIN; ++q) {
        if (!in[q]) return fail(EE_ERR_INVALID_ARG, "%s: input %d is null", name, q);
        vec_ok = vec_ok && aligned16(in[q]);
    }
    if (!out) return fail(EE_ERR_INVALID_ARG, "%s: out is null", name);
    const int64_t work = vec_ok ? ((n >> 2) + 1023) / 1024 : (n + 255) / 256;
    int64_t grid = work < 1 ? 1 : work;
    if (grid > 148 * 64) grid = 148 * 64;     // grid-stride beyond that
    ee::ew_kernel<NIN, F><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(i0, i1, i2, i3, i4, out, n, vec_ok, f);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, name);
    return EE_OK;
}

}  // namespace

template <int N, int R, int P>
static int launch_hfs(const ee::HfsArgs& a, cudaStream_t s) {
    using D_ = ee::HfsDims<N, R>;
    const size_t smem = (size_t)(D_::kTables + P * D_::kPlane) * sizeof(float);
    auto kernel = ee::hfs_kernel<N, R, P>;
    if (int rc = ensure_smem(kernel, smem)) return rc;
    // persistent CTAs: as many as can be resident (shared-memory limited), each loops over groups of P planes
    const int sms = sm_count();
    const int per_sm = (int)((size_t)(227 * 1024) / (smem + 1024));
    const int groups = (a.planes + P - 1) / P;
    int grid = sms * (per_sm < 1 ? 1 : per_sm);
    if (grid > groups) grid = groups;
    kernel<<<(unsigned)grid, 256, smem, s>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "ee_hfs_f32");
    return EE_OK;
}

// 64 px / r 8 on the tensor cores (ee_hfs_tc.cuh): persistent CTAs over pairs of planes, two per SM.  x is described to the
// TMA engine as a [planes * 64][64] fp32 matrix, box 32 x 128 (one K block of a pair of planes), 128 B swizzle.
static int launch_hfs_tc64(const ee::HfsArgs& a, cudaStream_t s) {
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static std::atomic<encode_fn> cached{nullptr};
    encode_fn fn = cached.load();
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) return fail(EE_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled is not available");
        fn = (encode_fn)p;
        cached.store(fn);
    }
    CUtensorMap map;
    const cuuint64_t dims[2] = {64u, (cuuint64_t)a.planes * 64u};
    const cuuint64_t strides[1] = {64u * sizeof(float)};
    const cuuint32_t box[2] = {32u, 128u};
    const cuuint32_t estr[2] = {1u, 1u};
    CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(a.x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(EE_ERR_UNSUPPORTED, "ee_hfs_tc_f32: cuTensorMapEncodeTiled failed (%d)", (int)r);
    CUtensorMap ymap;                               // y with a box of 32 x 32: one store per warp
    const cuuint32_t ybox[2] = {32u, 32u};
    r = fn(&ymap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, a.y, dims, strides, ybox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(EE_ERR_UNSUPPORTED, "ee_hfs_tc_f32: cuTensorMapEncodeTiled failed (%d)", (int)r);
    auto kernel = ee::hfs_tc::hfs_tc64_kernel<0>;
    if (int rc = ensure_smem(kernel, ee::hfs_tc::kSmem)) return rc;
    const int pairs = (a.planes + 1) / 2;
    int grid = 2 * sm_count();
    if (grid > pairs) grid = pairs;
    kernel<<<(unsigned)grid, ee::hfs_tc::kThreads, ee::hfs_tc::kSmem, s>>>(a, map, ymap);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "ee_hfs_tc_f32");
    return EE_OK;
}

template <int N, int R>
static int launch_hfs_rows(const ee::HfsArgs& a, cudaStream_t s) {
    const size_t smem = (size_t)ee::HfsRowsDims<N, R>::kFloats * sizeof(float);
    auto kernel = ee::hfs_rows_kernel<N, R>;
    if (int rc = ensure_smem(kernel, smem)) return rc;
    const int sms = sm_count();
    cudaError_t e;
    const int per_sm = (int)((size_t)(227 * 1024) / (smem + 1024));
    int grid = sms * (per_sm < 1 ? 1 : per_sm);
    if (grid > a.planes) grid = a.planes;
    kernel<<<(unsigned)grid, 256, smem, s>>>(a);
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "ee_hfs_f32");
    return EE_OK;
}

extern "C" {

int ee_edge_fwd_f32(const float* x, float* edge, int B, int C, int H, int W, const EEParams* p, void* stream) {
    return edge_forward(x, nullptr, nullptr, edge, B, C, H, W, p, 0.0f, false, stream);
}
int ee_edge_blend_fwd_f32(const float* x, const float* base, float* out, float* edge_or_null, int B, int C, int H,
                          int W, const EEParams* p, float w, void* stream) {
    return edge_forward(x, base, out, edge_or_null, B, C, H, W, p, w, true, stream);
}
int ee_edge_bwd_f32(const float* g_edge, const float* x, float* g_x, int B, int C, int H, int W, const EEParams* p,
                    void* stream) {
    return edge_backward(g_edge, x, nullptr, g_x, nullptr, B, C, H, W, p, 0.0f, false, stream);
}
int ee_edge_blend_bwd_f32(const float* g_out, const float* x, const float* base, float* g_x_or_null,
                          float* g_base_or_null, int B, int C, int H, int W, const EEParams* p, float w, void* stream) {
    return edge_backward(g_out, x, base, g_x_or_null, g_base_or_null, B, C, H, W, p, w, true, stream);
}
int ee_edge_fwd_strided_f32(const float* x, const EEStrides* xs, float* edge, const EEStrides* es, int B, int C, int H, int W,
                            const EEParams* p, void* stream) {
    return edge_forward(x, nullptr, nullptr, edge, B, C, H, W, p, 0.0f, false, stream, xs, nullptr, nullptr, es, true);
}
int ee_edge_bwd_strided_f32(const float* g_edge, const EEStrides* ges, const float* x, const EEStrides* xs, float* g_x,
                            const EEStrides* gxs, int B, int C, int H, int W, const EEParams* p, void* stream) {
    return edge_backward(g_edge, x, nullptr, g_x, nullptr, B, C, H, W, p, 0.0f, false, stream, ges, xs, nullptr, gxs, nullptr, true);
}
int ee_edge_blend_fwd_strided_f32(const float* x, const EEStrides* xs, const float* base, const EEStrides* bs, float* out,
                                  const EEStrides* os, float* edge_or_null, const EEStrides* es, int B, int C, int H, int W,
                                  const EEParams* p, float w, void* stream) {
    return edge_forward(x, base, out, edge_or_null, B, C, H, W, p, w, true, stream, xs, bs, os, es, true);
}
int ee_edge_blend_bwd_strided_f32(const float* g_out, const EEStrides* gs, const float* x, const EEStrides* xs, const float* base,
                                  const EEStrides* bs, float* g_x_or_null, const EEStrides* gxs, float* g_base_or_null,
                                  const EEStrides* gbs, int B, int C, int H, int W, const EEParams* p, float w, void* stream) {
    return edge_backward(g_out, x, base, g_x_or_null, g_base_or_null, B, C, H, W, p, w, true, stream, gs, xs, bs, gxs, gbs, true);
}
size_t ee_aux_bytes(int, int, int, int, int) { return 0; }

int ee_pgd_linf_step_f32(const float* x, const float* g, const float* x0, float* out, int64_t n, float alpha_signed,
                         float eps, float lo, float hi, void* stream) {
    return launch_ew<3>(x, g, x0, nullptr, nullptr, out, n, ee::FPgdLinf{alpha_signed, eps, lo, hi}, stream, "ee_pgd_linf_step_f32");
}
int ee_fgsm_step_f32(const float* x, const float* g, float* out, int64_t n, float alpha_signed, float lo, float hi,
                     void* stream) {
    return launch_ew<2>(x, g, nullptr, nullptr, nullptr, out, n, ee::FFgsm{alpha_signed, lo, hi}, stream, "ee_fgsm_step_f32");
}
int ee_cw_linf_step_f32(const float* adv, const float* g, const float* x, const float* min_x, const float* max_x,
                        float* out, int64_t n, float step, float magnitude, void* stream) {
    return launch_ew<5>(adv, g, x, min_x, max_x, out, n, ee::FCwLinf{step, magnitude}, stream, "ee_cw_linf_step_f32");
}
int ee_free_at_step_f32(float* delta, const float* g, const float* x0, float* x_adv, int64_t n, float alpha, float eps,
                        float lo, float hi, void* stream) {
    if (n < 0) return fail(EE_ERR_INVALID_ARG, "negative element count");
    if (n == 0) return EE_OK;
    if (!delta || !g || (x_adv && !x0)) return fail(EE_ERR_INVALID_ARG, "ee_free_at_step_f32: null pointer");
    const int vec_ok = aligned16(delta) && aligned16(g) && aligned16(x0) && aligned16(x_adv);
    const int64_t work = vec_ok ? ((n >> 2) + 1023) / 1024 : (n + 255) / 256;
    int64_t grid = work < 1 ? 1 : work;
    if (grid > 148 * 64) grid = 148 * 64;
    ee::free_at_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(delta, g, x0, x_adv, n, vec_ok, alpha, eps, lo, hi);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "ee_free_at_step_f32");
    return EE_OK;
}
int ee_add_clamp_f32(const float* x, const float* noise, float* out, int64_t n, float lo, float hi, void* stream) {
    return launch_ew<2>(x, noise, nullptr, nullptr, nullptr, out, n, ee::FAddClamp{lo, hi}, stream, "ee_add_clamp_f32");
}
int ee_avmixup_mix_f32(const float* x_adv, const float* inputs, const double* weight, float* out, int B, int64_t n_per,
                       float gamma, void* stream) {
    if (B < 0 || n_per < 0) return fail(EE_ERR_INVALID_ARG, "negative size");
    if (B == 0 || n_per == 0) return EE_OK;
    if (!x_adv || !inputs || !weight || !out) return fail(EE_ERR_INVALID_ARG, "ee_avmixup_mix_f32: null pointer");
    const int64_t n = (int64_t)B * n_per;
    const int vec_ok = (n_per % 4 == 0) && aligned16(x_adv) && aligned16(inputs) && aligned16(out);
    const int64_t work = ((vec_ok ? n / 4 : n) + 255) / 256;
    const unsigned grid = (unsigned)(work < 1 ? 1 : (work > 148 * 32 ? 148 * 32 : work));
    ee::avmixup_mix_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x_adv, inputs, weight, out, n, n_per, vec_ok, gamma);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "ee_avmixup_mix_f32");
    return EE_OK;
}
int ee_pgd_l2_step_f32(const float* x, const float* g, const float* x0, float* out, int B, int64_t n_per, float step,
                       float eps, void* stream) {
    if (B < 0 || n_per < 0) return fail(EE_ERR_INVALID_ARG, "negative size");
    if (B == 0 || n_per == 0) return EE_OK;
    if (!x || !g || !x0 || !out) return fail(EE_ERR_INVALID_ARG, "ee_pgd_l2_step_f32: null pointer");
    if (out == x) return fail(EE_ERR_INVALID_ARG, "ee_pgd_l2_step_f32: out must not alias x");
    cudaStream_t s = (cudaStream_t)stream;
    ee::L2Plan pl;
    // one pass (16 B/element): a cluster of K CTAs keeps the sample on chip between the two norms (ee_pgd_l2.cuh)
    if (g_staging.load() != 1 && aligned16(x) && aligned16(g) && aligned16(x0) && aligned16(out) && ee::pgd_l2_plan(n_per, pl) &&
        (long long)B * pl.K <= 0x7fffffffLL) {
        auto kernel = ee::pgd_l2_cluster_kernel;
        if (int rc = ensure_smem(kernel, pl.smem)) return rc;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)((long long)B * pl.K));
        cfg.blockDim = dim3((unsigned)pl.threads);
        cfg.dynamicSmemBytes = pl.smem;
        cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)pl.K; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, x, g, x0, out, n_per, pl.slice4, pl.K, pl.stage_x0, step, eps);
        if (e != cudaSuccess) return cuda_fail(e, "ee_pgd_l2_step_f32");
        return EE_OK;
    }
    // any size / alignment: three passes, one CTA per sample (a different, equally fixed, reduction order)
    ee::pgd_l2_kernel<<<(unsigned)B, 1024, 0, s>>>(x, g, x0, out, n_per, step, eps);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "ee_pgd_l2_step_f32");
    return EE_OK;
}

static int gf_blend(bool bwd, const float* g_out, const float* edge, const float* base, float* out, float* g_edge, float* g_base,
                    int B, int C, int H, int W, const float* gauss, float w, void* stream) {
    const char* name = bwd ? "ee_gf_blend_bwd_f32" : "ee_gf_blend_fwd_f32";
    if (B < 0 || C <= 0 || H <= 0 || W <= 0) return fail(EE_ERR_INVALID_ARG, "%s: bad shape", name);
    if (B == 0) return EE_OK;
    if (!edge || !base || !gauss || (bwd ? !g_out : !out)) return fail(EE_ERR_INVALID_ARG, "%s: null pointer", name);
    if (bwd && !g_edge && !g_base) return EE_OK;
    if (!(gauss[0] == gauss[2] && gauss[0] == gauss[6] && gauss[0] == gauss[8] && gauss[1] == gauss[3] && gauss[1] == gauss[5] &&
          gauss[1] == gauss[7]))
        return fail(EE_ERR_UNSUPPORTED, "%s: gauss[9] lacks the corner/edge/centre symmetry of get_gaussian_kernel(3,..)", name);
    if (B > 65535) return fail(EE_ERR_TOO_LARGE, "%s: more than 65535 images per call", name);
    ee::GfArgs a;
    a.edge = edge; a.base = base; a.g_out = g_out; a.out = out; a.g_edge = g_edge; a.g_base = g_base;
    a.B = B; a.C = C; a.H = H; a.W = W; a.c0 = gauss[0]; a.c1 = gauss[1]; a.c2 = gauss[4]; a.w = w;
    const dim3 grid((unsigned)((W + ee::kGfTW - 1) / ee::kGfTW), (unsigned)((H + ee::kGfTH - 1) / ee::kGfTH), (unsigned)B);
    if (bwd) ee::gf_blend_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    else ee::gf_blend_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, name);
    return EE_OK;
}
int ee_gf_blend_fwd_f32(const float* edge, const float* base, float* out, int B, int C, int H, int W, const float gauss[9],
                        float w, void* stream) {
    return gf_blend(false, nullptr, edge, base, out, nullptr, nullptr, B, C, H, W, gauss, w, stream);
}
int ee_gf_blend_bwd_f32(const float* g_out, const float* edge, const float* base, float* g_edge_or_null, float* g_base_or_null,
                        int B, int C, int H, int W, const float gauss[9], float w, void* stream) {
    return gf_blend(true, g_out, edge, base, nullptr, g_edge_or_null, g_base_or_null, B, C, H, W, gauss, w, stream);
}

int ee_edge_pgd_iteration_f32(const float* x, const float* base, const float* g_out, const float* x0, float* out_or_null,
                              float* g_x, float* g_base_or_null, float* x_next, int B, int C, int H, int W, const EEParams* p,
                              float w, float alpha_signed, float eps, void* stream) {
    if (!g_x || !x_next || !x0) return fail(EE_ERR_INVALID_ARG, "ee_edge_pgd_iteration_f32: null pointer");
    if (x_next == x) return fail(EE_ERR_INVALID_ARG, "ee_edge_pgd_iteration_f32: x_next must not alias x");
    int rc = EE_OK;
    if (out_or_null) rc = edge_forward(x, base, out_or_null, nullptr, B, C, H, W, p, w, true, stream);
    if (rc) return rc;
    rc = edge_backward(g_out, x, base, g_x, g_base_or_null, B, C, H, W, p, w, true, stream);
    if (rc) return rc;
    return launch_ew<3>(x, g_x, x0, nullptr, nullptr, x_next, (int64_t)B * C * H * W, ee::FPgdLinf{alpha_signed, eps, 0.0f, 1.0f},
                        stream, "ee_edge_pgd_iteration_f32");
}

int ee_to_compare_fwd_f32(const float* in, float* out, int64_t n, float thr, void* stream) {
    return launch_ew<1>(in, nullptr, nullptr, nullptr, nullptr, out, n, ee::FToCompareFwd{thr}, stream, "ee_to_compare_fwd_f32");
}
int ee_to_compare_bwd_f32(const float* g, const float* in, float* out, int64_t n, float thr, void* stream) {
    return launch_ew<2>(g, in, nullptr, nullptr, nullptr, out, n, ee::FToCompareBwd{thr}, stream, "ee_to_compare_bwd_f32");
}
int ee_to_eq_fwd_f32(const float* in, float* out, int64_t n, void* stream) {
    return launch_ew<1>(in, nullptr, nullptr, nullptr, nullptr, out, n, ee::FToEqFwd{}, stream, "ee_to_eq_fwd_f32");
}
int ee_to_eq_bwd_f32(const float* g, const float* in, float* out, int64_t n, void* stream) {
    return launch_ew<2>(g, in, nullptr, nullptr, nullptr, out, n, ee::FToEqBwd{}, stream, "ee_to_eq_bwd_f32");
}
int ee_safe_sign_fwd_f32(const float* in, float* out, int64_t n, void* stream) {
    return launch_ew<1>(in, nullptr, nullptr, nullptr, nullptr, out, n, ee::FSafeSignFwd{}, stream, "ee_safe_sign_fwd_f32");
}
int ee_safe_sign_bwd_f32(const float* g, const float* in, float* out, int64_t n, void* stream) {
    return launch_ew<2>(g, in, nullptr, nullptr, nullptr, out, n, ee::FSafeSignBwd{}, stream, "ee_safe_sign_bwd_f32");
}

static int add_square(bool bwd, const float* g, const float* x, const float* stripe, const float* table, float* out, int B,
                      int C, int H, int W, int n_sq, float eps, void* stream) {
    const char* name = bwd ? "ee_add_square_bwd_f32" : "ee_add_square_fwd_f32";
    if (B < 0 || C <= 0 || H <= 0 || W <= 0 || n_sq < 0) return fail(EE_ERR_INVALID_ARG, "%s: bad shape", name);
    if (B == 0) return EE_OK;
    if (!x || !stripe || !out || (bwd && !g) || (n_sq > 0 && !table)) return fail(EE_ERR_INVALID_ARG, "%s: null pointer", name);
    ee::SquareArgs a;
    a.x = x; a.stripe = stripe; a.table = table; a.g = g; a.out = out;
    a.C = C; a.H = H; a.W = W; a.n_sq = n_sq; a.n = (int64_t)B * C * H * W; a.eps = eps;
    const bool vec = (W % 4 == 0) && aligned16(x) && aligned16(stripe) && aligned16(out) && aligned16(g);
    const int64_t work = ((vec ? a.n / 4 : a.n) + 255) / 256;
    const unsigned grid = (unsigned)(work < 1 ? 1 : (work > 148 * 32 ? 148 * 32 : work));
    cudaStream_t s = (cudaStream_t)stream;
    if (bwd) { if (vec) ee::add_square_kernel<true, 4><<<grid, 256, 0, s>>>(a); else ee::add_square_kernel<true, 1><<<grid, 256, 0, s>>>(a); }
    else     { if (vec) ee::add_square_kernel<false, 4><<<grid, 256, 0, s>>>(a); else ee::add_square_kernel<false, 1><<<grid, 256, 0, s>>>(a); }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, name);
    return EE_OK;
}
int ee_add_square_fwd_f32(const float* x, const float* stripe, const float* table, float* out, int B, int C, int H, int W,
                          int n_sq, float eps, void* stream) {
    return add_square(false, nullptr, x, stripe, table, out, B, C, H, W, n_sq, eps, stream);
}
int ee_add_square_bwd_f32(const float* g, const float* x, const float* stripe, const float* table, float* g_x, int B, int C,
                          int H, int W, int n_sq, float eps, void* stream) {
    return add_square(true, g, x, stripe, table, g_x, B, C, H, W, n_sq, eps, stream);
}

int ee_hfs_supported(int N, int r) {
    // the reference's configurations: MNIST 28 / 4, Tiny-ImageNet 64 / 8, ImageNet 224 / 16, fast-AT schedule 128 / 12 and 288 / 18
    return (N == 64 && r == 8) || (N == 28 && r == 4) || (N == 32 && r == 8) || (N == 224 && r == 16) || (N == 128 && r == 12) ||
           (N == 288 && r == 18);
}
int ee_hfs_f32(const float* x, float* y, const float* add_or_null, int planes, int N, int r, const float* cb, const float* rb,
               const float* w, float gamma, void* stream) {
    if (planes < 0) return fail(EE_ERR_INVALID_ARG, "ee_hfs_f32: negative plane count");
    if (planes == 0) return EE_OK;
    if (!x || !y || !cb || !rb || !w) return fail(EE_ERR_INVALID_ARG, "ee_hfs_f32: null pointer");
    if (x == y) return fail(EE_ERR_INVALID_ARG, "ee_hfs_f32: y must not alias x");
    if (!aligned16(x) || !aligned16(y) || !aligned16(add_or_null) || !aligned16(cb) || !aligned16(rb) || !aligned16(w))
        return fail(EE_ERR_INVALID_ARG, "ee_hfs_f32: pointers must be 16-byte aligned");
    ee::HfsArgs a;
    a.x = x; a.y = y; a.add = add_or_null; a.cb = cb; a.rb = rb; a.w = w; a.gamma = gamma; a.planes = planes;
    cudaStream_t s = (cudaStream_t)stream;
    if (N == 64 && r == 8) return launch_hfs<64, 8, 4>(a, s);
    if (N == 28 && r == 4) return launch_hfs<28, 4, 16>(a, s);
    if (N == 32 && r == 8) return launch_hfs<32, 8, 8>(a, s);
    if (N == 224 && r == 16) return launch_hfs_rows<224, 16>(a, s);      // ImageNet: row-blocked, one plane per CTA
    if (N == 128 && r == 12) return launch_hfs<128, 12, 2>(a, s);          // 2 whole planes per CTA still fit (208 KB)
    if (N == 288 && r == 18) return launch_hfs_rows<288, 18>(a, s);
    return fail(EE_ERR_UNSUPPORTED, "ee_hfs_f32: no kernel for a %d x %d plane with radius %d (ee_hfs_supported)", N, N, r);
}

int ee_hfs_tc_supported(int N, int r) { return N == 64 && r == 8; }
int ee_hfs_tc_f32(const float* x, float* y, const float* add_or_null, int planes, int N, int r, const float* cb, const float* rb,
                  const float* w, float gamma, void* stream) {
    if (planes < 0) return fail(EE_ERR_INVALID_ARG, "ee_hfs_tc_f32: negative plane count");
    if (!ee_hfs_tc_supported(N, r))
        return fail(EE_ERR_UNSUPPORTED, "ee_hfs_tc_f32: the tensor-core kernel exists for 64 x 64 planes with radius 8 only (got %d / %d)", N, r);
    if (planes == 0) return EE_OK;
    if (!x || !y || !cb || !rb || !w) return fail(EE_ERR_INVALID_ARG, "ee_hfs_tc_f32: null pointer");
    if (x == y) return fail(EE_ERR_INVALID_ARG, "ee_hfs_tc_f32: y must not alias x");
    if (!aligned16(x) || !aligned16(y) || !aligned16(add_or_null) || !aligned16(cb) || !aligned16(rb) || !aligned16(w))
        return fail(EE_ERR_INVALID_ARG, "ee_hfs_tc_f32: pointers must be 16-byte aligned");
    ee::HfsArgs a;
    a.x = x; a.y = y; a.add = add_or_null; a.cb = cb; a.rb = rb; a.w = w; a.gamma = gamma; a.planes = planes;
    return launch_hfs_tc64(a, (cudaStream_t)stream);
}

const char* ee_last_error(void) { return g_err; }
int ee_version(void) { return EE_VERSION; }
int ee_set_tuning(int strip_rows_fwd, int strip_rows_bwd, int staging) {
    g_th_fwd.store(strip_rows_fwd < 0 ? 0 : strip_rows_fwd);
    g_th_bwd.store(strip_rows_bwd < 0 ? 0 : strip_rows_bwd);
    g_staging.store(staging);
    return EE_OK;
}

}  // extern "C"
#endif  // EE_HAS(1)
