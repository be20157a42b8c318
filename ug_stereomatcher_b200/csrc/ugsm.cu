// ugsm.cu -- host-side level driver and the C-ABI (include/ugsm.h) over the sm_100a kernels.
//
// Mirrors the orchestration of the reference's MatchGPULib (src/gpu_matcher/MatchGPULib.cpp):
//   match()               MGL:303-403    -> ugsm_match_full
//   matchStackPyramid()   MGL:534-700    -> ugsm_match_foveated
//   CreatePyramidFromImage MGL:1033-1125 -> build_pyramid()
//   CreateFoveatedPyramid MGL:1128-1190  -> crops in prepare_levels()
//   matching()            MGL:1196-1318  -> run_levels()
//   matchlevel()          MGL:1662-2489  -> run_level()  (K1 + K2 per iteration)
//   subsampleDisp / foveatedsubsampleDisp MGL:1526-1655 -> K3
// but keeps everything resident on the device: no host round trips inside a pair, one
// stream per handle, no global device state, no cudaDeviceReset.
#include "../../include/ugsm.h"

#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "ugsm_kernels.cuh"
#include "ugsm_k1_stream.cuh"
#include "ugsm_k2_stream.cuh"

using namespace ugsm;

#define SCALE_D 1.41421356 /* MatchLib_common.h:15 */

namespace {

// Error text of calls that have no handle (ugsm_create, ugsm_host_alloc): per thread, so that handles created from
// several threads (ugsm_batch: one handle per worker) neither race on it nor read a dangling pointer.
thread_local char g_create_error[512] = "";

inline int align_up(int v, int a) { return (v + a - 1) / a * a; }
inline int div_up(int a, int b) { return (a + b - 1) / b; }

struct LevelBuf {
    int w = 0, h = 0, pitch = 0;     // natural pyramid dims
    int mw = 0, mh = 0, mpitch = 0;  // dims the matcher runs at
    int crop_x = 0, crop_y = 0;
    bool cropped = false;
    float* L[3] = {nullptr, nullptr, nullptr};
    float* R[3] = {nullptr, nullptr, nullptr};
    float* Lm[3] = {nullptr, nullptr, nullptr};  // == L/R unless cropped
    float* Rm[3] = {nullptr, nullptr, nullptr};
    float* U[3] = {nullptr, nullptr, nullptr};
    float* st[3] = {nullptr, nullptr, nullptr};  // H, V, C
    int iters = 0, smooth = 0;
};

}  // namespace

struct ugsm_handle {
    ugsm_config cfg;
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string err;
    Taps taps;
    int sm_count = 148;

    // shape-dependent state
    int rows = 0, cols = 0, mode = -1;
    ugsm_dims dims;
    std::vector<LevelBuf> lv;
    void* arena = nullptr;
    size_t arena_bytes = 0;
    float* tmp[3] = {nullptr, nullptr, nullptr};  // K1 output / K2 input
    uint8_t* d_rgb[2] = {nullptr, nullptr};       // staging for the host-buffer entry points
    size_t d_rgb_bytes = 0;
    float* d_out[3] = {nullptr, nullptr, nullptr};  // dense output staging
    size_t d_out_floats = 0;

    cudaEvent_t ev[4 + UGSM_MAX_LEVELS] = {};
    // per-launch events of the UGSM_TIMED_LEVELS finest levels: 3 per iteration (before K1, between, after K2)
    static constexpr int KEV_MAX = 3 * 64;
    cudaEvent_t kev[KEV_MAX] = {};
    int kev_used = 0;
    int kev_level[64] = {};
    ugsm_timing timing;
    int launches = 0;
    bool k_attr_set = false;
    int k1_nw = 2;
    int k1_stream = 1;        // large levels use the streaming strip march k1_stream (UGSM_K1_STREAM=0: the tiled k1_correlate)
    int k1s_slots = 0;        // co-resident k1_stream CTAs on the device (occupancy x SM count)
    int k1s_min_seg = 8;      // shortest row segment of a strip (UGSM_K1S_MIN_SEG)
    int* d_flags = nullptr;   // per level: nonzero when the six matcher planes lie in [1, 256] (k1_stream FAST march)
    int k1_fast = 1;          // UGSM_K1_FAST=0: always take the guarded GENERIC march
    int k1_tma = 1;           // left image / U / C rows of k1_stream arrive as TMA tiles (UGSM_K1_TMA=0: cp.async)
    void* tma_encode = nullptr;  // cuTensorMapEncodeTiled, fetched through cudaGetDriverEntryPoint
    int k2_stream = 1;        // large levels use the row pipeline k2_stream (UGSM_K2_STREAM=0: the tiled k2_smooth_box)
    int k2s_slots[K2S_MAX_PASSES + 1] = {};  // co-resident k2_stream CTAs on the device per pass count
    int k2s_min_seg = 24;     // shortest row segment of a strip (UGSM_K2S_MIN_SEG)
    long k2s_min_pixels = 400000;   // k2_stream runs levels of at least this many pixels (UGSM_K2S_MIN_PIXELS) ...
    int k2s_max_passes = 10;        // ... and at most this many smoothing passes (UGSM_K2S_MAX_PASSES)
    int k2_tma = 1;                 // k2_stream's loader uses TMA tile loads (UGSM_K2_TMA)
    int pdl = 1;                    // programmatic dependent launch of the coarse-level kernels (UGSM_PDL)
    int k2s_edge_extra = 2;         // extra row segments of the edge strips, whose steps are slower (UGSM_K2S_EDGE_EXTRA)
    long k2_tiny_pixels = 20000;     // levels below this use k2_tiny_tile-sided k2_smooth_box_small tiles (UGSM_K2_TINY_PIXELS)
    int k2_tiny_tile = 8;           // (UGSM_K2_TINY_TILE)
    long k1_tiny_pixels = 20000;    // levels below this use 2-row k1_correlate_small tiles (UGSM_K1_TINY_PIXELS)
    long k1_small_pixels = 200000;  // levels below this use k1_correlate_small (UGSM_K1_SMALL_PIXELS)
    long k2_small_pixels = 400000;  // levels below this use k2_smooth_box_small (UGSM_K2_SMALL_PIXELS)
    int k0_rows_dense = 16, k0_rows_dec = 8;  // output rows per warp of k0_march5 (UGSM_K0_ROWS_DENSE / _DEC)
    // CUDA graph of the per-pair launch sequence (pyramid levels .. last level), valid for the current shape
    cudaGraphExec_t graph_exec = nullptr;
    int graph_launches = 0;
    // same for the row-band path: pyramid + levels [levels-1 .. band_split] + upsample into band_split-1
    cudaGraphExec_t band_graph = nullptr;
    int band_graph_split = -1, band_graph_launches = 0;
    // hierarchicalDisparity composites: 3 planes per level below the fovea level (natural dims), lazily allocated
    void* comp_arena = nullptr;
    float* comp[UGSM_MAX_LEVELS][3] = {};
    // point cloud (K5): what the last match left on the device, and grow-only staging for the host entry points
    const uint8_t* last_rgbL = nullptr;
    size_t last_step = 0;
    int last_kind = -1;  // 0 full, 1 foveated, 2 foveated + composite
    void* cloud_buf = nullptr;
    size_t cloud_bytes = 0;
    // pageable host buffers (malloc'd planes of the reference API): pinned staging + copy threads
    int host_threads = 8;  // UGSM_HOST_THREADS; < 2 = plain cudaMemcpyAsync from/to the caller's memory
    char* pin_in = nullptr;
    size_t pin_in_bytes = 0;
    char* pin_out = nullptr;
    size_t pin_out_bytes = 0;
    std::vector<cudaEvent_t> chunk_ev;
    cudaEvent_t cloud_ev[2] = {};
    float cloud_ms = 0.f;
};

namespace {

int fail(ugsm_handle* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) {
        try {
            h->err = buf;
        } catch (...) {  // out of memory while recording an error: keep the code, drop the text
        }
    } else {
        snprintf(g_create_error, sizeof g_create_error, "%s", buf);
    }
    return code;
}

#define CK(call)                                                                                             \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess)                                                                               \
            return fail(h, UGSM_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                        __LINE__);                                                                           \
    } while (0)

// MGL:761-775 and MGL:344-348, evaluated in float exactly as the host code does.
void make_taps(Taps& t) {
    float k[5];
    k[0] = 0.0816475;
    k[1] = 0.218507;
    k[2] = 0.303281;
    k[3] = 0.218507;
    k[4] = 0.0816475;
    float kern = 0;
    for (int i = 0; i < 5; i++) kern = kern + k[i];
    for (int i = 0; i < 5; i++) t.g[i] = k[i] / kern;
    t.a[0] = 0.0;
    t.a[1] = 0.3333;
    t.a[2] = 0.3333;
    t.a[3] = 0.3333;
    t.a[4] = 0.0;
}

int level_iters(const ugsm_config& c, int i) {  // MGL:1741 with level = MAX_LEVEL-1-i
    int mi = (i > 5) ? c.iters_cutoff : ((i + 1) * 2);
    if (c.max_iters > 0 && mi > c.max_iters) mi = c.max_iters;
    return mi;
}
int level_smooth(const ugsm_config& c, int i) {  // MGL:2257-2261
    return (i < 2) ? c.smooth_passes_fine : c.smooth_passes;
}
// MGL:2299-2306: evaluated after iteration m, applies from iteration m+1.
float threshold_after(int mi, int m, float T) {
    float threshold = T;
    if (m % 2 == 0) {
        if ((mi / 2 - m / 2) < 7) {
            threshold = ((mi / 2 - m / 2) - 1) * ((1 - 0.1) / (mi / 2 - 1.0)) + 0.1;
        } else {
            threshold = 1.0;
        }
    }
    return threshold;
}

int validate_cfg(const ugsm_config& c) {
    if (c.levels < 1 || c.levels > UGSM_MAX_LEVELS) return 0;
    if (c.iters_cutoff < 1 || c.smooth_passes < 0 || c.smooth_passes_fine < 0) return 0;
    if (c.smooth_passes > 16 || c.smooth_passes_fine > 16) return 0;
    return 1;
}

int compute_dims(const ugsm_config& c, int rows, int cols, int mode, ugsm_dims* d) {
    if (!validate_cfg(c) || rows < 1 || cols < 1 || (mode != 0 && mode != 1)) return UGSM_ERR_INVALID;
    memset(d, 0, sizeof *d);
    d->levels = c.levels;
    d->width[0] = cols;
    d->height[0] = rows;
    for (int i = 0; i + 1 < c.levels; i++) {  // MGL:1224-1228
        d->width[i + 1] = d->width[i] / SCALE_D;
        d->height[i + 1] = d->height[i] / SCALE_D;
    }
    if (d->width[c.levels - 1] < 1 || d->height[c.levels - 1] < 1) return UGSM_ERR_INVALID;
    int fl = c.fovea_level;
    if (mode == 1 && (fl < 2 || fl > c.levels)) return UGSM_ERR_INVALID;
    if (fl >= 1 && fl <= c.levels) {
        d->fov_width = d->width[fl - 1];
        d->fov_height = d->height[fl - 1];
        d->stack_rows = fl * d->fov_height;
    }
    for (int i = 0; i < c.levels; i++) {
        bool crop = (mode == 1 && i < fl - 1);  // MGL:1230-1240
        d->match_width[i] = crop ? d->fov_width : d->width[i];
        d->match_height[i] = crop ? d->fov_height : d->height[i];
        d->iters[i] = level_iters(c, i);
        d->smooth[i] = level_smooth(c, i);
        d->pixel_iterations += (int64_t)d->match_width[i] * d->match_height[i] * d->iters[i];
    }
    return UGSM_OK;
}

// ------------------------------------------------------------------ kernel launchers
constexpr int K1_TH = 16;
constexpr int K1S_TH = 4;       // rows of a k1_correlate_small tile (two half-strips of 2 rows: 6 march steps)
constexpr int K1S_TH_TINY = 2;  // ... on the tiniest levels (5 march steps, twice the CTAs)
// K0: the register-marching kernels win on large planes (HBM-bound), the tiled one on small planes (latency-bound)
constexpr long K0_MARCH_DENSE_PIXELS = 900000;  // output pixels
constexpr long K0_MARCH_DEC_PIXELS = 200000;
constexpr int K2_MAX_BH = 72;  // region rows (tile + 2 halos): 2*3*72*128*4 B = 216 KB of shared memory (227 KB max)

size_t k2_smem_bytes(int bh) { return (size_t)2 * 3 * K2_BW * bh * sizeof(float); }

int set_kernel_attrs(ugsm_handle* h) {
    if (h->k_attr_set) return UGSM_OK;
    CK(cudaFuncSetAttribute(k1_correlate<4, K1_TH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)sizeof(K1Smem<4, K1_TH>)));
    CK(cudaFuncSetAttribute(k1_correlate<2, K1_TH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)sizeof(K1Smem<2, K1_TH>)));
    CK(cudaFuncSetAttribute(k1_correlate<2, K1_TH>, cudaFuncAttributePreferredSharedMemoryCarveout,
                            cudaSharedmemCarveoutMaxShared));
    if (const char* e = getenv("UGSM_K1_NW")) h->k1_nw = atoi(e);  // tuning override: warps per K1 CTA (2 or 4)
    if (const char* e = getenv("UGSM_K1_STREAM")) h->k1_stream = atoi(e);
    if (const char* e = getenv("UGSM_K1_FAST")) h->k1_fast = atoi(e);
    if (const char* e = getenv("UGSM_K1S_MIN_SEG")) h->k1s_min_seg = atoi(e) > 0 ? atoi(e) : h->k1s_min_seg;
    if (const char* e = getenv("UGSM_K1_TMA")) h->k1_tma = atoi(e);
    {
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &h->tma_encode, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            h->tma_encode = nullptr;  // then k1_stream<false> (all cp.async) runs
        cudaGetLastError();
        CK(cudaFuncSetAttribute(k1_stream<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        CK(cudaFuncSetAttribute(k1_stream<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        int nb = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k1_stream<true>, K1S_NW * 32, K1S_NW * sizeof(K1Ring)));
        h->k1s_slots = std::max(1, nb) * h->sm_count;
        if (const char* e = getenv("UGSM_K1S_SLOTS")) h->k1s_slots = atoi(e) > 0 ? atoi(e) : h->k1s_slots;
    }
    if (const char* e = getenv("UGSM_K1_SMALL_PIXELS")) h->k1_small_pixels = atol(e);
    if (const char* e = getenv("UGSM_K1_TINY_PIXELS")) h->k1_tiny_pixels = atol(e);
    if (const char* e = getenv("UGSM_K2_TINY_PIXELS")) h->k2_tiny_pixels = atol(e);
    if (const char* e = getenv("UGSM_K2_TINY_TILE")) h->k2_tiny_tile = std::min(K2S_T, std::max(2, atoi(e)));
    if (const char* e = getenv("UGSM_K2_SMALL_PIXELS")) h->k2_small_pixels = atol(e);
    if (const char* e = getenv("UGSM_K0_ROWS_DENSE")) h->k0_rows_dense = atoi(e) > 0 ? atoi(e) : h->k0_rows_dense;
    if (const char* e = getenv("UGSM_K0_ROWS_DEC")) h->k0_rows_dec = atoi(e) > 0 ? atoi(e) : h->k0_rows_dec;
    CK(cudaFuncSetAttribute(k2_smooth_box, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)k2_smem_bytes(K2_MAX_BH)));
    if (const char* e = getenv("UGSM_K2_STREAM")) h->k2_stream = atoi(e);
    if (const char* e = getenv("UGSM_K2S_MIN_SEG")) h->k2s_min_seg = atoi(e) > 0 ? atoi(e) : h->k2s_min_seg;
    if (const char* e = getenv("UGSM_K2S_MIN_PIXELS")) h->k2s_min_pixels = atol(e);
    if (const char* e = getenv("UGSM_K2S_EDGE_EXTRA")) h->k2s_edge_extra = std::max(0, atoi(e));
    if (const char* e = getenv("UGSM_PDL")) h->pdl = atoi(e);
    if (const char* e = getenv("UGSM_K2_TMA")) h->k2_tma = atoi(e);
    if (const char* e = getenv("UGSM_K2S_MAX_PASSES")) h->k2s_max_passes = std::min(K2S_MAX_PASSES, std::max(0, atoi(e)));
    CK(cudaFuncSetAttribute(k2_stream<K2S_NT_SMALL, K2S_MINB_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)k2s_smem_bytes(K2S_NT_SMALL / 32 - 2)));
    CK(cudaFuncSetAttribute(k2_stream<K2S_NT_SMALL, K2S_MINB_SMALL>, cudaFuncAttributePreferredSharedMemoryCarveout,
                            cudaSharedmemCarveoutMaxShared));
    CK(cudaFuncSetAttribute(k2_stream<K2S_NT_LARGE, K2S_MINB_LARGE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)k2s_smem_bytes(K2S_MAX_PASSES)));
    CK(cudaFuncSetAttribute(k2_stream<K2S_NT_LARGE, K2S_MINB_LARGE>, cudaFuncAttributePreferredSharedMemoryCarveout,
                            cudaSharedmemCarveoutMaxShared));
    for (int sps = 0; sps <= K2S_MAX_PASSES; sps++) {
        int nb = 0;
        if ((sps + 2) * 32 <= K2S_NT_SMALL)
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k2_stream<K2S_NT_SMALL, K2S_MINB_SMALL>, (sps + 2) * 32,
                                                             k2s_smem_bytes(sps)));
        else
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k2_stream<K2S_NT_LARGE, K2S_MINB_LARGE>, (sps + 2) * 32,
                                                             k2s_smem_bytes(sps)));
        h->k2s_slots[sps] = std::max(1, nb) * h->sm_count;
    }
    if (const char* e = getenv("UGSM_K2S_SLOTS"))
        for (int sps = 0; sps <= K2S_MAX_PASSES; sps++) h->k2s_slots[sps] = atoi(e) > 0 ? atoi(e) : h->k2s_slots[sps];
    h->k_attr_set = true;
    return UGSM_OK;
}

// 2-D f32 tensor map of a w x hh plane (row pitch `pitch` floats): tiles of box_w x box_h, zero fill outside the
// image.  False when the plane cannot be described (base not 16-byte aligned, pitch not a multiple of 4 floats).
bool make_plane_map(ugsm_handle* h, CUtensorMap* tm, const float* plane, int w, int hh, int pitch, int box_w, int box_h) {
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    if (!h->tma_encode || (reinterpret_cast<uintptr_t>(plane) & 15) || (pitch & 3)) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)w, (cuuint64_t)hh};
    const cuuint64_t strides[1] = {(cuuint64_t)pitch * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h};
    const cuuint32_t estr[2] = {1, 1};
    return reinterpret_cast<encode_fn>(h->tma_encode)(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(plane),
                                                      dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Launch of a launch-bound kernel as a programmatic dependent of its predecessor in the stream (pdl_wait /
// pdl_trigger in ugsm_kernels.cuh); UGSM_PDL=0 restores ordinary launches.  No dynamic shared memory.
template <typename... KArgs, typename... Args>
void launch_pdl(ugsm_handle* h, void (*kernel)(KArgs...), dim3 grid, dim3 block, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = 0;
    cfg.stream = h->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    // Inside a graph capture the attribute would become a programmatic edge; measured, graph replay is faster with
    // ordinary kernel-to-kernel edges (16 MP pair 13.8 vs 14.2 ms), so only plain stream launches use it
    // (use_graph = 0: 15.0 -> 14.6 ms).
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (h->pdl) cudaStreamIsCapturing(h->stream, &cs);
    cfg.numAttrs = (h->pdl && cs == cudaStreamCaptureStatusNone) ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

void launch_k1(ugsm_handle* h, const K1Params& p, const int* fast_flag = nullptr) {
    // Large levels: streaming strip march (k1_stream): one warp per 30-column strip and row segment; the
    // segment length is chosen so that all CTAs of the launch are co-resident (one wave, no tail).
    // Small levels: 30 x 8 tiles with one warp per colour channel, so that the few tiles there spread
    // over more SMs and the dependent chain per launch is short.
    const int band_rows = p.y_end - p.y_begin;
    const bool big = (long)p.w * p.h >= h->k1_small_pixels || band_rows != p.h;  // a row band always uses a band-aware kernel
    if (big && h->k1_stream) {
        K1StreamParams sp;
        sp.k = p;
        sp.fast_flag = h->k1_fast ? fast_flag : nullptr;
        const int strips = div_up(p.w, 30 * K1S_NW);
        int nseg = std::max(1, h->k1s_slots / strips);
        nseg = std::min(nseg, std::max(1, band_rows / h->k1s_min_seg));
        sp.seg_rows = div_up(band_rows, nseg);
        dim3 grid(strips, div_up(band_rows, sp.seg_rows));
        bool tma = h->k1_tma != 0;
        for (int k = 0; k < 3 && tma; k++) tma = make_plane_map(h, &sp.tm[k], p.L[k], p.w, p.h, p.pitch, K1S_LP, K1S_CH);
        if (tma)
            k1_stream<true><<<grid, K1S_NW * 32, K1S_NW * sizeof(K1Ring), h->stream>>>(sp);
        else
            k1_stream<false><<<grid, K1S_NW * 32, K1S_NW * sizeof(K1Ring), h->stream>>>(sp);
    } else if (big && h->k1_nw == 2) {
        dim3 grid(div_up(p.w, 60), div_up(band_rows, K1_TH));
        k1_correlate<2, K1_TH><<<grid, 64, sizeof(K1Smem<2, K1_TH>), h->stream>>>(p);
    } else if (big) {
        dim3 grid(div_up(p.w, 120), div_up(band_rows, K1_TH));
        k1_correlate<4, K1_TH><<<grid, 128, sizeof(K1Smem<4, K1_TH>), h->stream>>>(p);
    } else {
        // latency-bound: the march pays a 4-step prologue per tile, so short tiles on more SMs are faster
        // (8 -> 4 rows: 16 MP pair 13.33 -> 13.16 ms, 640x480 3.33 -> 2.85 ms; 2 rows only pay on the tiniest levels)
        if ((long)p.w * p.h < h->k1_tiny_pixels) {
            dim3 grid(div_up(p.w, 30), div_up(p.h, K1S_TH_TINY));
            launch_pdl(h, k1_correlate_small<K1S_TH_TINY>, grid, dim3(256), p);
        } else {
            dim3 grid(div_up(p.w, 30), div_up(p.h, K1S_TH));
            launch_pdl(h, k1_correlate_small<K1S_TH>, grid, dim3(256), p);
        }
    }
    h->launches++;
}

void launch_k2(ugsm_handle* h, const float* Hi, const float* Vi, const float* Ci, float* Ho, float* Vo, float* Co,
               int w, int hh, int pitch, int passes, int y_begin = 0, int y_end = -1) {
    if (y_end < 0) y_end = hh;
    const bool banded = (y_begin != 0 || y_end != hh);
    K2Params p;
    p.y_begin = y_begin;
    p.y_end = y_end;
    p.Hi = Hi;
    p.Vi = Vi;
    p.Ci = Ci;
    p.Ho = Ho;
    p.Vo = Vo;
    p.Co = Co;
    p.w = w;
    p.h = hh;
    p.pitch = pitch;
    p.passes = passes;
    p.a1 = h->taps.a[1];
    const int R = passes + 1;
    if (!banded && (long)w * hh < h->k2_small_pixels && K2S_T + 2 * R <= 32) {
        // small levels: 16x16 tiles, one thread per region cell (latency-oriented)
        // latency-bound like K1-small: on the tiniest levels smaller tiles (fewer warps at the per-pass barrier, more
        // SMs busy) are faster although the halo is recomputed more often
        const int tt = ((long)w * hh < h->k2_tiny_pixels) ? h->k2_tiny_tile : K2S_T;
        p.th = tt;
        p.hpad = R;
        const int side = tt + 2 * R;
        dim3 grid(div_up(w, tt), div_up(hh, tt));
        launch_pdl(h, k2_smooth_box_small, grid, dim3(align_up(side * side, 32)), p);
        h->launches++;
        return;
    }
    p.hpad = align_up(R, 4);
    if (h->k2_stream && passes <= h->k2s_max_passes && (long)w * hh >= h->k2s_min_pixels) {
        // levels of 400 k pixels and more, S <= 10: row pipeline over the passes (k2_stream), row segments sized for
        // one co-resident wave (two CTAs per SM)
        K2StreamParams sp;
        sp.k = p;
        sp.k.th = 0;
        const int strips = div_up(w, K2_BW - 2 * p.hpad);
        const int brows = y_end - y_begin;
        // one co-resident wave: interior strips x nseg segments plus the (at most 3) edge strips' extra segment
        int nseg = std::max(1, (h->k2s_slots[passes] - 3 * h->k2s_edge_extra) / strips);
        nseg = std::min(nseg, std::max(1, brows / h->k2s_min_seg));
        sp.nseg = nseg;
        sp.edge_extra = h->k2s_edge_extra;
        {   // input rows as TMA tiles: descriptors over pitch x hh floats (pad columns read as the cp.async loader does)
            const float* src[3] = {Hi, Vi, Ci};
            bool tma = h->k2_tma != 0;
            for (int k = 0; k < 3 && tma; k++) tma = make_plane_map(h, &sp.tm[k], src[k], pitch, hh, pitch, K2_BW, 1);
            sp.use_tma = tma ? 1 : 0;
        }
        dim3 grid(strips, nseg + sp.edge_extra);
        if ((passes + 2) * 32 <= K2S_NT_SMALL)
            k2_stream<K2S_NT_SMALL, K2S_MINB_SMALL><<<grid, (passes + 2) * 32, k2s_smem_bytes(passes), h->stream>>>(sp);
        else
            k2_stream<K2S_NT_LARGE, K2S_MINB_LARGE><<<grid, (passes + 2) * 32, k2s_smem_bytes(passes), h->stream>>>(sp);
        h->launches++;
        return;
    }
    const int twu = K2_BW - 2 * p.hpad;
    // Tile height: as tall as shared memory allows on big levels (less halo redundancy); shorter on
    // small levels so that there are at least ~2 CTAs per SM.
    int th = K2_MAX_BH - 2 * R;
    const int cols = div_up(w, twu);
    const int brows = y_end - y_begin;
    while (th > 8 && (long)cols * div_up(brows, th) < 2L * h->sm_count) th = (th + 1) / 2;
    if (th > brows) th = brows;
    p.th = th;
    dim3 grid(cols, div_up(brows, th));
    k2_smooth_box<<<grid, K2_NWARP * 32, k2_smem_bytes(th + 2 * R), h->stream>>>(p);
    h->launches++;
}

void launch_k3(ugsm_handle* h, float* const src[3], int w, int hh, int spitch, int crop_x, int crop_y,
               float* const dst[3], int ow, int oh, int dpitch, int oy_begin = 0, int oy_end = -1) {
    if (oy_end < 0) oy_end = oh;
    dim3 block(32, 8), grid(div_up(ow, 32), div_up(oy_end - oy_begin, 8));
    launch_pdl(h, k3_upsample, grid, block, (const float*)src[0], (const float*)src[1], (const float*)src[2], w, hh, spitch,
               crop_x, crop_y, dst[0], dst[1], dst[2], ow, oy_end, dpitch, oy_begin);
    h->launches++;
}

// The register-marching K0 kernels use aligned vector accesses.
bool planes_aligned(const K0Planes& pl, int n, int spitch, int dpitch) {
    if ((spitch & 3) || (dpitch & 3)) return false;
    for (int k = 0; k < n; k++)
        if ((reinterpret_cast<uintptr_t>(pl.src[k]) & 15) || (reinterpret_cast<uintptr_t>(pl.dst[k]) & 15)) return false;
    return true;
}

// flag[0] = nonzero iff every value of the six matcher planes of a level lies in [1, 256]: the precondition of
// k1_stream's FAST march (ugsm_k1_stream.cuh, k1s_fast_bounds).  Stays on the device: no host round trip.
void launch_range_flag(ugsm_handle* h, const float* const* L, const float* const* R, int w, int hh, int pitch,
                       int* flag) {
    RangePlanes pl;
    for (int k = 0; k < 3; k++) {
        pl.p[k] = L[k];
        pl.p[3 + k] = R[k];
    }
    cudaMemsetAsync(flag, 1, sizeof(int), h->stream);
    dim3 grid(std::min(div_up(w, 1024), 8), std::min(hh, 512), 6);
    k_range_flag<<<grid, 256, 0, h->stream>>>(pl, w, hh, pitch, 1.0f, 256.0f, flag);
    h->launches++;
}

// U = colT(rowT(L^2)) for `n` planes of equal dims in one launch.
void launch_energy(ugsm_handle* h, const float* const* L, float* const* U, int n, int w, int hh, int pitch) {
    K0Planes pl;
    for (int k = 0; k < 6; k++) {
        pl.src[k] = L[k < n ? k : 0];
        pl.dst[k] = U[k < n ? k : 0];
    }
    if ((long)w * hh >= K0_MARCH_DENSE_PIXELS && planes_aligned(pl, n, pitch, pitch)) {
        const int rpw = h->k0_rows_dense;
        dim3 grid(div_up(w, 128), div_up(hh, 8 * rpw), n);
        k0_march5<0, true, true><<<grid, dim3(32, 8), 0, h->stream>>>(pl, w, hh, pitch, w, hh, pitch, rpw, h->taps);
    } else {
        dim3 grid(div_up(w, 32), div_up(hh, 32), n);
        k0_sep5<0, true, true><<<grid, 256, 0, h->stream>>>(pl, w, hh, pitch, w, hh, pitch, 1.0f, h->taps);
    }
    h->launches++;
}

void launch_unpack(ugsm_handle* h, const uint8_t* rgb, size_t step, int w, int hh, float* const pl[3], int pitch) {
    dim3 grid(div_up(w, 256), hh);
    k0_unpack_rgb8<<<grid, 256, 0, h->stream>>>(rgb, step, w, hh, pl[0], pl[1], pl[2], pitch);
    h->launches++;
}

// Fused blur + decimate of `n` planes of equal dims in one launch (mode 0: sqrt2 NN, mode 1: odd samples).
void launch_blur_decimate(ugsm_handle* h, int mode, const float* const* src, float* const* dst, int n, int w, int hh,
                          int spitch, int w2, int h2, int dpitch) {
    K0Planes pl;
    for (int k = 0; k < 6; k++) {
        pl.src[k] = src[k < n ? k : 0];
        pl.dst[k] = dst[k < n ? k : 0];
    }
    dim3 grid(div_up(w2, 32), div_up(h2, 32), n);
    if (mode == 0) {
        float s = SCALE_D;  // MGL:1083
        k0_sep5<2, false, false><<<grid, 256, 0, h->stream>>>(pl, w, hh, spitch, w2, h2, dpitch, s, h->taps);
    } else if ((long)w2 * h2 >= K0_MARCH_DEC_PIXELS && planes_aligned(pl, n, spitch, dpitch)) {
        // scalefactors = (int)(SCALE*SCALE + .5) = 2 (MGL:1090): the odd samples (2x+1, 2y+1)
        const int rpw = h->k0_rows_dec;
        dim3 mgrid(div_up(w2, 64), div_up(h2, 8 * rpw), n);
        k0_march5<1, false, false><<<mgrid, dim3(32, 8), 0, h->stream>>>(pl, w, hh, spitch, w2, h2, dpitch, rpw,
                                                                          h->taps);
    } else {
        float s = 0.000 + (int)(SCALE_D * SCALE_D + 0.5);  // MGL:1090
        k0_sep5<1, false, false><<<grid, 256, 0, h->stream>>>(pl, w, hh, spitch, w2, h2, dpitch, s, h->taps);
    }
    h->launches++;
}

void launch_copy3(ugsm_handle* h, const float* s0, const float* s1, const float* s2, int spitch, float* d0, float* d1,
                  float* d2, int dpitch, int w, int hh) {
    dim3 grid(div_up(w, 256), hh);
    k_copy3<<<grid, 256, 0, h->stream>>>(s0, s1, s2, spitch, d0, d1, d2, dpitch, w, hh);
    h->launches++;
}

// ------------------------------------------------------------------ per-shape buffers
void release_shape(ugsm_handle* h) {
    if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
    h->graph_exec = nullptr;
    if (h->band_graph) cudaGraphExecDestroy(h->band_graph);
    h->band_graph = nullptr;
    h->band_graph_split = -1;
    if (h->comp_arena) cudaFree(h->comp_arena);
    h->comp_arena = nullptr;
    if (h->arena) cudaFree(h->arena);
    h->arena = nullptr;
    h->arena_bytes = 0;
    h->lv.clear();
    h->rows = h->cols = 0;
    h->mode = -1;
    h->last_kind = -1;
}

int prepare_shape(ugsm_handle* h, int rows, int cols, int mode) {
    if (h->arena && h->rows == rows && h->cols == cols && h->mode == mode) return UGSM_OK;
    release_shape(h);
    ugsm_dims d;
    int rc = compute_dims(h->cfg, rows, cols, mode, &d);
    if (rc != UGSM_OK) return fail(h, rc, "invalid shape/config: rows=%d cols=%d mode=%d levels=%d", rows, cols, mode,
                                   h->cfg.levels);
    const int levels = h->cfg.levels;
    std::vector<LevelBuf> lv(levels);
    size_t total = 0;
    auto plane_bytes = [](int pitch, int hh) { return ((size_t)pitch * hh * sizeof(float) + 255) / 256 * 256; };
    size_t max_match = 0;
    for (int i = 0; i < levels; i++) {
        LevelBuf& b = lv[i];
        b.w = d.width[i];
        b.h = d.height[i];
        b.pitch = align_up(b.w, 32);
        b.mw = d.match_width[i];
        b.mh = d.match_height[i];
        b.mpitch = align_up(b.mw, 32);
        b.cropped = (b.mw != b.w || b.mh != b.h);
        if (b.cropped) {  // MGL:1143-1176: l = x/2 - fovW/2, u = y/2 - fovH/2
            b.crop_x = b.w / 2 - b.mw / 2;
            b.crop_y = b.h / 2 - b.mh / 2;
        }
        b.iters = d.iters[i];
        b.smooth = d.smooth[i];
        total += 6 * plane_bytes(b.pitch, b.h);
        if (b.cropped) total += 6 * plane_bytes(b.mpitch, b.mh);
        total += 6 * plane_bytes(b.mpitch, b.mh);  // U[3] + state[3]
        size_t mb = plane_bytes(b.mpitch, b.mh);
        if (mb > max_match) max_match = mb;
    }
    total += 3 * max_match;
    CK(cudaMalloc(&h->arena, total));
    h->arena_bytes = total;
    char* cur = static_cast<char*>(h->arena);
    auto take = [&](size_t bytes) {
        float* p = reinterpret_cast<float*>(cur);
        cur += bytes;
        return p;
    };
    for (int i = 0; i < levels; i++) {
        LevelBuf& b = lv[i];
        for (int k = 0; k < 3; k++) b.L[k] = take(plane_bytes(b.pitch, b.h));
        for (int k = 0; k < 3; k++) b.R[k] = take(plane_bytes(b.pitch, b.h));
        for (int k = 0; k < 3; k++) {
            b.Lm[k] = b.cropped ? take(plane_bytes(b.mpitch, b.mh)) : b.L[k];
            b.Rm[k] = b.cropped ? take(plane_bytes(b.mpitch, b.mh)) : b.R[k];
        }
        for (int k = 0; k < 3; k++) b.U[k] = take(plane_bytes(b.mpitch, b.mh));
        for (int k = 0; k < 3; k++) b.st[k] = take(plane_bytes(b.mpitch, b.mh));
    }
    for (int k = 0; k < 3; k++) h->tmp[k] = take(max_match);
    h->lv.swap(lv);
    h->dims = d;
    h->rows = rows;
    h->cols = cols;
    h->mode = mode;
    return UGSM_OK;
}

// MGL:1033-1125 for `nimg` images (left [+ right]): 3*nimg planes per launch.
void unpack_images(ugsm_handle* h, const uint8_t* d_rgbL, const uint8_t* d_rgbR, size_t step) {
    launch_unpack(h, d_rgbL, step, h->lv[0].w, h->lv[0].h, h->lv[0].L, h->lv[0].pitch);
    if (d_rgbR) launch_unpack(h, d_rgbR, step, h->lv[0].w, h->lv[0].h, h->lv[0].R, h->lv[0].pitch);
}

void build_pyramid(ugsm_handle* h, int nimg) {
    const int levels = h->cfg.levels;
    auto planes = [&](int i, const float** out) {
        for (int k = 0; k < 3; k++) {
            out[k] = h->lv[i].L[k];
            out[3 + k] = h->lv[i].R[k];
        }
    };
    for (int i = 0; i < levels; i++) {
        const LevelBuf& s = h->lv[i];
        const float* src[6];
        const float* dstc[6];
        planes(i, src);
        if (i == 0 && levels > 1) {
            planes(1, dstc);
            launch_blur_decimate(h, 0, src, (float* const*)dstc, 3 * nimg, s.w, s.h, s.pitch, h->lv[1].w, h->lv[1].h,
                                 h->lv[1].pitch);
        }
        if (i + 2 < levels) {
            planes(i + 2, dstc);
            launch_blur_decimate(h, 1, src, (float* const*)dstc, 3 * nimg, s.w, s.h, s.pitch, h->lv[i + 2].w,
                                 h->lv[i + 2].h, h->lv[i + 2].pitch);
        }
    }
}

// Foveated crops (MGL:1171-1185) and the per-level left-image energy U.
void prepare_levels(ugsm_handle* h) {
    for (auto& b : h->lv) {
        if (b.cropped) {
            dim3 grid(div_up(b.mw, 256), b.mh);
            for (int k = 0; k < 3; k++) {
                k_crop<<<grid, 256, 0, h->stream>>>(b.L[k], b.pitch, b.crop_x, b.crop_y, b.Lm[k], b.mw, b.mh,
                                                    b.mpitch);
                k_crop<<<grid, 256, 0, h->stream>>>(b.R[k], b.pitch, b.crop_x, b.crop_y, b.Rm[k], b.mw, b.mh,
                                                    b.mpitch);
                h->launches += 2;
            }
        }
        launch_energy(h, b.Lm, b.U, 3, b.mw, b.mh, b.mpitch);
        const int i = (int)(&b - &h->lv[0]);
        if (h->k1_stream && h->k1_fast && (long)b.mw * b.mh >= h->k1_small_pixels)
            launch_range_flag(h, b.Lm, b.Rm, b.mw, b.mh, b.mpitch, h->d_flags + i);
    }
}

// MGL:1662-2489: `iters` iterations of K1 (correlate/fit/blend) + K2 (smooth/box).
void run_level(ugsm_handle* h, float* const L[3], float* const R[3], float* const U[3], float* const st[3],
               float* const tmp[3], int w, int hh, int pitch, int iters, int passes, bool coarsest,
               int timed_level = -1, const int* fast_flag = nullptr) {
    float T = 1.0f;  // MGL:1673
    for (int m = 1; m <= iters; m++) {
        int slot = -1;
        if (timed_level >= 0 && h->kev_used < 64) {
            slot = h->kev_used++;
            h->kev_level[slot] = timed_level;
            cudaEventRecord(h->kev[3 * slot], h->stream);
        }
        K1Params p;
        for (int k = 0; k < 3; k++) {
            p.L[k] = L[k];
            p.R[k] = R[k];
            p.U[k] = U[k];
        }
        p.H = st[0];
        p.V = st[1];
        p.C = st[2];
        p.Hn = tmp[0];
        p.Vn = tmp[1];
        p.Cn = tmp[2];
        p.w = w;
        p.h = hh;
        p.pitch = pitch;
        p.y_begin = 0;
        p.y_end = hh;
        p.T = T;
        p.first_coarsest = (coarsest && m == 1) ? 1 : 0;
        p.dbg_srcx = nullptr;
        p.dbg_srcy = nullptr;
        p.dbg_bh = nullptr;
        p.dbg_bv = nullptr;
        p.tp = h->taps;
        launch_k1(h, p, fast_flag);
        if (slot >= 0) cudaEventRecord(h->kev[3 * slot + 1], h->stream);
        launch_k2(h, tmp[0], tmp[1], tmp[2], st[0], st[1], st[2], w, hh, pitch, passes);
        if (slot >= 0) cudaEventRecord(h->kev[3 * slot + 2], h->stream);
        T = threshold_after(iters, m, T);
    }
}

// MGL:1196-1318 coarse-to-fine loop over levels [hi .. lo].  `events` = record the per-level and
// per-launch timing events (not possible while capturing a graph: those cannot be timed).
int run_levels(ugsm_handle* h, int hi, int lo, bool events) {
    const int levels = h->cfg.levels;
    const int fl = h->cfg.fovea_level;
    for (int i = hi; i >= lo; i--) {
        LevelBuf& b = h->lv[i];
        if (i == levels - 1) {
            // zero start (the reference reads uninitialised malloc here, MGL:1247)
            for (int k = 0; k < 3; k++)
                CK(cudaMemsetAsync(b.st[k], 0, (size_t)b.mpitch * b.mh * sizeof(float), h->stream));
        } else {
            LevelBuf& c = h->lv[i + 1];
            if (h->mode == 0 || i + 1 >= fl) {  // MGL:1279 / 1283-1286
                launch_k3(h, c.st, c.mw, c.mh, c.mpitch, 0, 0, b.st, b.mw, b.mh, b.mpitch);
            } else {  // MGL:1290 + 1612-1615: upsample to level fl-2 dims, centre crop
                int w2 = h->dims.width[fl - 2], h2 = h->dims.height[fl - 2];
                launch_k3(h, c.st, c.mw, c.mh, c.mpitch, w2 / 2 - c.mw / 2, h2 / 2 - c.mh / 2, b.st, b.mw, b.mh,
                          b.mpitch);
            }
        }
        run_level(h, b.Lm, b.Rm, b.U, b.st, h->tmp, b.mw, b.mh, b.mpitch, b.iters, b.smooth, i == levels - 1,
                  (events && i < UGSM_TIMED_LEVELS) ? i : -1, h->d_flags + i);
        if (events) CK(cudaEventRecord(h->ev[4 + i], h->stream));
    }
    return UGSM_OK;
}

// Whole device-resident pair: pyramid of both images, crops/energy, coarse-to-fine loop.
int run_pair(ugsm_handle* h, const uint8_t* d_rgbL, const uint8_t* d_rgbR, int rows, int cols, size_t step,
             int mode) {
    if (!d_rgbL || !d_rgbR) return fail(h, UGSM_ERR_INVALID, "null image pointer");
    if (step < (size_t)cols * 3) return fail(h, UGSM_ERR_INVALID, "step %zu < cols*3", step);
    CK(cudaSetDevice(h->device));
    int rc = set_kernel_attrs(h);
    if (rc) return rc;
    rc = prepare_shape(h, rows, cols, mode);
    if (rc) return rc;
    h->launches = 0;
    h->kev_used = 0;
    CK(cudaEventRecord(h->ev[0], h->stream));
    unpack_images(h, d_rgbL, d_rgbR, step);  // reads the caller's buffers: stays outside the graph
    const int levels = h->cfg.levels;
    if (!h->cfg.use_graph || levels < 2) {
        build_pyramid(h, 2);
        prepare_levels(h);
        CK(cudaEventRecord(h->ev[1], h->stream));
        rc = run_levels(h, levels - 1, 0, true);
        if (rc) return rc;
    } else {
        // Graph: pyramid + crops/energy + the whole coarse-to-fine loop (hundreds of short launches, where the
        // ~4 us launch floor matters).  use_graph == 2 (instrumentation: bench.py's roofline region, tools/) keeps
        // level 0 outside the graph so that its K1 / K2 launches get timing events; production (1) does not pay
        // for that.
        const int graph_lo = (h->cfg.use_graph == 2) ? 1 : 0;
        if (!h->graph_exec) {
            // Capture once per shape: every buffer the sequence touches belongs to the handle's arena.
            const int l0 = h->launches;
            cudaGraph_t g = nullptr;
            CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed));
            build_pyramid(h, 2);
            prepare_levels(h);
            rc = run_levels(h, levels - 1, graph_lo, false);
            cudaError_t e = cudaStreamEndCapture(h->stream, &g);
            if (rc || e != cudaSuccess || !g) {
                if (g) cudaGraphDestroy(g);
                return rc ? rc : fail(h, UGSM_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
            }
            e = cudaGraphInstantiate(&h->graph_exec, g, 0);
            cudaGraphDestroy(g);
            if (e != cudaSuccess) {
                h->graph_exec = nullptr;
                return fail(h, UGSM_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
            }
            h->graph_launches = h->launches - l0;
            h->launches = l0;
        }
        CK(cudaGraphLaunch(h->graph_exec, h->stream));
        h->launches += h->graph_launches;
        if (graph_lo == 1) {
            CK(cudaEventRecord(h->ev[4 + 1], h->stream));
            rc = run_levels(h, 0, 0, true);
            if (rc) return rc;
        }
    }
    CK(cudaEventRecord(h->ev[2], h->stream));
    CK(cudaGetLastError());
    h->last_rgbL = d_rgbL;
    h->last_step = step;
    h->last_kind = mode;
    return UGSM_OK;
}

int finish_timing(ugsm_handle* h) {
    CK(cudaEventSynchronize(h->ev[2]));
    ugsm_timing& t = h->timing;
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, h->ev[0], h->ev[2]));
    t.total_ms = ms;
    t.launches = h->launches;
    t.pyramid_ms = t.levels_ms = 0.f;
    for (int i = 0; i < UGSM_MAX_LEVELS; i++) t.level_ms[i] = t.k1_ms[i] = t.k2_ms[i] = 0.f;
    const int levels = h->cfg.levels;
    const bool graph = h->cfg.use_graph && levels >= 2;  // then at most level 0 has inner events
    if (graph && h->cfg.use_graph != 2) return UGSM_OK;  // everything replayed from the graph: only the total
    float v = 0;
    if (!graph) {
        CK(cudaEventElapsedTime(&v, h->ev[0], h->ev[1]));
        t.pyramid_ms = v;
        CK(cudaEventElapsedTime(&v, h->ev[1], h->ev[2]));
        t.levels_ms = v;
    }
    for (int i = graph ? 0 : levels - 1; i >= 0; i--) {
        cudaEvent_t prev = (i == levels - 1) ? h->ev[1] : h->ev[4 + i + 1];
        CK(cudaEventElapsedTime(&v, prev, h->ev[4 + i]));
        t.level_ms[i] = v;
    }
    for (int s = 0; s < h->kev_used; s++) {
        CK(cudaEventElapsedTime(&v, h->kev[3 * s], h->kev[3 * s + 1]));
        t.k1_ms[h->kev_level[s]] += v;
        CK(cudaEventElapsedTime(&v, h->kev[3 * s + 1], h->kev[3 * s + 2]));
        t.k2_ms[h->kev_level[s]] += v;
    }
    return UGSM_OK;
}

int ensure_staging(ugsm_handle* h, size_t rgb_bytes, size_t out_floats) {
    if (rgb_bytes > h->d_rgb_bytes) {
        for (int k = 0; k < 2; k++) {
            if (h->d_rgb[k]) cudaFree(h->d_rgb[k]);
            h->d_rgb[k] = nullptr;
        }
        h->d_rgb_bytes = 0;
        for (int k = 0; k < 2; k++) CK(cudaMalloc(&h->d_rgb[k], rgb_bytes));
        h->d_rgb_bytes = rgb_bytes;
    }
    if (out_floats > h->d_out_floats) {
        for (int k = 0; k < 3; k++) {
            if (h->d_out[k]) cudaFree(h->d_out[k]);
            h->d_out[k] = nullptr;
        }
        h->d_out_floats = 0;
        for (int k = 0; k < 3; k++) CK(cudaMalloc(&h->d_out[k], out_floats * sizeof(float)));
        h->d_out_floats = out_floats;
    }
    return UGSM_OK;
}

// Write the results of the last run_pair into dense device planes.
int emit_full(ugsm_handle* h, float* dH, float* dV, float* dC) {
    LevelBuf& b = h->lv[0];
    launch_copy3(h, dH ? b.st[0] : nullptr, dV ? b.st[1] : nullptr, dC ? b.st[2] : nullptr, b.mpitch, dH, dV, dC, b.mw,
                 b.mw, b.mh);
    CK(cudaGetLastError());
    return UGSM_OK;
}
int emit_stack(ugsm_handle* h, float* sH, float* sV, float* sC) {
    const int fl = h->cfg.fovea_level;
    for (int i = 0; i < fl; i++) {  // level-major rows, UG_GPU_matcher.cpp:293-303
        LevelBuf& b = h->lv[i];
        size_t off = (size_t)i * b.mw * b.mh;
        launch_copy3(h, b.st[0], b.st[1], b.st[2], b.mpitch, sH + off, sV + off, sC + off, b.mw, b.mw, b.mh);
    }
    CK(cudaGetLastError());
    return UGSM_OK;
}

// MGL:2589-2701 on the state the foveated run left in the level buffers; result in comp[0].
int run_composite(ugsm_handle* h) {
    const int fl = h->cfg.fovea_level;
    if (h->mode != 1 || fl < 2) return fail(h, UGSM_ERR_INVALID, "composite needs a foveated run");
    if (!h->comp_arena) {
        size_t total = 0;
        for (int i = 0; i < fl - 1; i++) total += 3 * (((size_t)h->lv[i].pitch * h->lv[i].h * sizeof(float) + 255) / 256 * 256);
        CK(cudaMalloc(&h->comp_arena, total));
        char* cur = static_cast<char*>(h->comp_arena);
        for (int i = 0; i < fl - 1; i++)
            for (int k = 0; k < 3; k++) {
                h->comp[i][k] = reinterpret_cast<float*>(cur);
                cur += ((size_t)h->lv[i].pitch * h->lv[i].h * sizeof(float) + 255) / 256 * 256;
            }
    }
    const float levelscale = SCALE_D;  // MGL:2638 (double) -> float parameter of partsubsampleDispGPU
    for (int level = fl - 1; level > 0; level--) {
        LevelBuf& c = h->lv[level];      // coarser level
        LevelBuf& b = h->lv[level - 1];  // this level: fovea state b.st (b.mw x b.mh), natural dims b.w x b.h
        K4Planes pl;
        int pw, ph, pp;
        if (level == fl - 1) {
            pw = c.mw, ph = c.mh, pp = c.mpitch;
        } else {
            pw = c.w, ph = c.h, pp = c.pitch;
        }
        for (int k = 0; k < 3; k++) {
            pl.prev[k] = (level == fl - 1) ? c.st[k] : h->comp[level][k];
            pl.fov[k] = b.st[k];
            pl.dst[k] = h->comp[level - 1][k];
        }
        dim3 block(32, 8), grid(div_up(b.w, 32), div_up(b.h, 8));
        k4_composite<<<grid, block, 0, h->stream>>>(pl, pw, ph, pp, b.mw, b.mh, b.mpitch, b.crop_x, b.crop_y, b.w, b.h,
                                                    b.pitch, levelscale);
        h->launches++;
    }
    CK(cudaGetLastError());
    h->last_kind = 2;
    return UGSM_OK;
}
int emit_composite(ugsm_handle* h, float* dH, float* dV, float* dC) {
    LevelBuf& b = h->lv[0];
    launch_copy3(h, h->comp[0][0], h->comp[0][1], h->comp[0][2], b.pitch, dH, dV, dC, b.w, b.w, b.h);
    CK(cudaGetLastError());
    return UGSM_OK;
}

struct TmpDev {  // scoped device allocations for the ugsm_stage_* entry points
    std::vector<void*> ptrs;
    ~TmpDev() {
        for (void* p : ptrs) cudaFree(p);
    }
    template <class T>
    T* alloc(size_t n) {
        void* p = nullptr;
        if (cudaMalloc(&p, n * sizeof(T)) != cudaSuccess) return nullptr;
        ptrs.push_back(p);
        return static_cast<T*>(p);
    }
};

// dense host plane -> pitched device plane and back
cudaError_t up2d(float* d, int pitch, const float* hsrc, int w, int hh, cudaStream_t s) {
    return cudaMemcpy2DAsync(d, (size_t)pitch * 4, hsrc, (size_t)w * 4, (size_t)w * 4, hh, cudaMemcpyHostToDevice, s);
}
cudaError_t down2d(float* hdst, const float* d, int pitch, int w, int hh, cudaStream_t s) {
    return cudaMemcpy2DAsync(hdst, (size_t)w * 4, d, (size_t)pitch * 4, (size_t)w * 4, hh, cudaMemcpyDeviceToHost, s);
}

// ---- K5: point cloud (PC = src/pointcloud/getPointCloud.cpp) ----
// PC:432-484 with destLevel = 0: margin of the fovea of stack level src_level inside the full image.
void cloud_margins(int im_rows, int im_cols, int fovea_level, int src_level, int* left, int* upper) {
    int w = im_cols, hh = im_rows;
    for (int k = 0; k < (fovea_level - 1) - src_level; k++) {
        w = (int)(w / SCALE_D);
        hh = (int)(hh / SCALE_D);
    }
    *left = im_cols / 2 - w / 2;
    *upper = im_rows / 2 - hh / 2;
}

int cloud_events(ugsm_handle* h) {
    for (auto& ev : h->cloud_ev)
        if (!ev) CK(cudaEventCreate(&ev));
    return UGSM_OK;
}

int ensure_cloud_staging(ugsm_handle* h, size_t bytes) {
    if (bytes <= h->cloud_bytes) return UGSM_OK;
    if (h->cloud_buf) cudaFree(h->cloud_buf);
    h->cloud_buf = nullptr;
    h->cloud_bytes = 0;
    CK(cudaMalloc(&h->cloud_buf, bytes));
    h->cloud_bytes = bytes;
    return UGSM_OK;
}

// Launch K5 on device-resident planes.  src_level < 0: full-resolution planes; otherwise the planes are one
// level of a fovea stack belonging to an im_rows x im_cols image.
int launch_cloud(ugsm_handle* h, const ugsm_camera_pair* cam, const float* dX, const float* dY, int pitch, int rows,
                 int cols, int src_level, int im_rows, int im_cols, const uint8_t* d_img, size_t step, int bgr,
                 int sampling, float* d_points, int stride, int64_t* npoints, float* d_range = nullptr) {
    // d_range: the range map (z of every pixel, dense rows x cols) instead of a cloud; no image, no points
    if (!cam || !dX || !dY || (!d_range && (!d_img || !d_points))) return fail(h, UGSM_ERR_INVALID, "null pointer argument");
    if (rows < 1 || cols < 1 || sampling < 1) return fail(h, UGSM_ERR_INVALID, "bad rows/cols/sampling");
    if (!d_range && stride != 4 && stride != 8) return fail(h, UGSM_ERR_INVALID, "point stride must be 4 or 8 floats");
    if (!d_range && (reinterpret_cast<uintptr_t>(d_points) & 15) != 0)
        return fail(h, UGSM_ERR_INVALID, "points not 16-byte aligned");
    K5Params p;
    memcpy(p.P1, cam->P1, sizeof p.P1);
    memcpy(p.P2, cam->P2, sizeof p.P2);
    p.dX = dX, p.dY = dY, p.pitch = pitch, p.rows = rows, p.cols = cols;
    p.img = d_img, p.step = step, p.bgr = bgr ? 1 : 0, p.sampling = sampling;
    p.ncol = div_up(cols, sampling), p.nrow = div_up(rows, sampling);
    p.points = d_points, p.stride = stride;
    p.range = d_range;
    p.foveated = 0, p.left = p.upper = 0, p.scale = 1.0f;
    if (src_level >= 0) {
        if (src_level >= h->cfg.fovea_level)
            return fail(h, UGSM_ERR_INVALID, "src_level %d outside the stack", src_level);
        p.foveated = 1;
        cloud_margins(im_rows, im_cols, h->cfg.fovea_level, src_level, &p.left, &p.upper);
        // PC:385-388: pow(float, float) is the float overload in C++11 -> libm powf, as on the reference host
        p.scale = powf((float)1.41421356237309504880, (float)src_level);
    }
    int rc = cloud_events(h);
    if (rc) return rc;
    dim3 block(K5_T, 8), grid(div_up(p.ncol, K5_T), div_up(p.nrow, K5_T));
    CK(cudaEventRecord(h->cloud_ev[0], h->stream));
    k5_cloud<<<grid, block, 0, h->stream>>>(p);
    CK(cudaEventRecord(h->cloud_ev[1], h->stream));
    CK(cudaGetLastError());
    h->launches++;
    if (npoints) *npoints = (int64_t)p.ncol * p.nrow;
    return UGSM_OK;
}

int finish_cloud(ugsm_handle* h) {
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaEventElapsedTime(&h->cloud_ms, h->cloud_ev[0], h->cloud_ev[1]));
    return UGSM_OK;
}

// ---- pageable host buffers -------------------------------------------------------------------------
// The reference API hands out malloc'd planes (MGL:372-399: the node frees them), i.e. pageable memory:
// a plain cudaMemcpy of 96 MB in + 193 MB out costs ~50 ms at 16 MP, most of it first-touch page faults
// and the driver's single-threaded staging.  Pageable buffers therefore go through pinned staging
// buffers in chunks, with a few copy threads on the host side: user -> pinned copies run ahead of the
// H2D DMA; the output pages are pre-faulted while the GPU computes, and pinned -> user copies follow
// the D2H DMA chunk by chunk.
constexpr size_t HOST_CHUNK = (size_t)4 << 20;

bool is_pageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

struct HostXfer {  // one buffer of a staged transfer
    const void* src;  // H2D: caller's memory; D2H: device memory
    void* dst;        // H2D: device memory;   D2H: caller's memory
    size_t bytes;
};
struct HostChunk {
    const char* src;
    char* dst;
    char* pin;
    size_t len;
};

int ensure_pinned(ugsm_handle* h, size_t in_bytes, size_t out_bytes) {
    if (in_bytes > h->pin_in_bytes) {
        if (h->pin_in) cudaFreeHost(h->pin_in);
        h->pin_in = nullptr;
        h->pin_in_bytes = 0;
        CK(cudaHostAlloc((void**)&h->pin_in, in_bytes, cudaHostAllocDefault));
        h->pin_in_bytes = in_bytes;
    }
    if (out_bytes > h->pin_out_bytes) {
        if (h->pin_out) cudaFreeHost(h->pin_out);
        h->pin_out = nullptr;
        h->pin_out_bytes = 0;
        CK(cudaHostAlloc((void**)&h->pin_out, out_bytes, cudaHostAllocDefault));
        h->pin_out_bytes = out_bytes;
    }
    return UGSM_OK;
}

size_t staged_bytes(const HostXfer* x, int n) {
    size_t t = 0;
    for (int i = 0; i < n; i++) t += (x[i].bytes + 255) / 256 * 256;
    return t;
}

std::vector<HostChunk> make_chunks(const HostXfer* x, int n, char* pinned) {
    std::vector<HostChunk> ch;
    for (int i = 0; i < n; i++) {
        for (size_t off = 0; off < x[i].bytes; off += HOST_CHUNK)
            ch.push_back({(const char*)x[i].src + off, (char*)x[i].dst + off, pinned + off,
                          std::min(HOST_CHUNK, x[i].bytes - off)});
        pinned += (x[i].bytes + 255) / 256 * 256;
    }
    return ch;
}

// Start up to `want` copy threads; never throws (no exception may cross the C-ABI).  Returns how many run.
template <class F>
int start_threads(std::vector<std::thread>& pool, int want, F&& work) {
    try {
        pool.reserve(want);
        for (int t = 0; t < want; t++) pool.emplace_back(work);
    } catch (...) {
    }
    return (int)pool.size();
}

// caller (pageable) -> pinned -> device.  Copy threads fill chunks in any order; the calling thread enqueues
// the DMA of a chunk as soon as it is filled, in order.
int staged_h2d_impl(ugsm_handle* h, const HostXfer* x, int n);
int staged_d2h_impl(ugsm_handle* h, const HostXfer* x, int n);
int staged_h2d(ugsm_handle* h, const HostXfer* x, int n) {
    try {
        return staged_h2d_impl(h, x, n);
    } catch (...) {
        return fail(h, UGSM_ERR_NOMEM, "staged H2D: host allocation failed");
    }
}
int staged_d2h(ugsm_handle* h, const HostXfer* x, int n) {
    try {
        return staged_d2h_impl(h, x, n);
    } catch (...) {
        return fail(h, UGSM_ERR_NOMEM, "staged D2H: host allocation failed");
    }
}

int staged_h2d_impl(ugsm_handle* h, const HostXfer* x, int n) {
    const std::vector<HostChunk> ch = make_chunks(x, n, h->pin_in);
    const int total = (int)ch.size();
    std::unique_ptr<std::atomic<int>[]> done(new std::atomic<int>[total]);
    for (int c = 0; c < total; c++) done[c].store(0, std::memory_order_relaxed);
    std::atomic<int> next(0);
    auto work = [&]() {
        for (;;) {
            const int c = next.fetch_add(1);
            if (c >= total) break;
            memcpy(ch[c].pin, ch[c].src, ch[c].len);
            done[c].store(1, std::memory_order_release);
        }
    };
    std::vector<std::thread> pool;
    if (start_threads(pool, h->host_threads, work) == 0) work();  // no thread could be started: copy inline
    cudaError_t err = cudaSuccess;
    for (int c = 0; c < total; c++) {
        while (!done[c].load(std::memory_order_acquire)) std::this_thread::yield();
        if (err == cudaSuccess) err = cudaMemcpyAsync(ch[c].dst, ch[c].pin, ch[c].len, cudaMemcpyHostToDevice, h->stream);
    }
    for (auto& t : pool) t.join();
    if (err != cudaSuccess) return fail(h, UGSM_ERR_CUDA, "staged H2D failed: %s", cudaGetErrorString(err));
    return UGSM_OK;
}

// device -> pinned -> caller (pageable).  Called right after the producing kernels were enqueued: the copy
// threads first pre-fault the caller's pages (while the GPU is busy), then follow the D2H DMA chunk by chunk.
int staged_d2h_impl(ugsm_handle* h, const HostXfer* x, int n) {
    const std::vector<HostChunk> ch = make_chunks(x, n, h->pin_out);
    const int total = (int)ch.size();
    while ((int)h->chunk_ev.size() < total) {
        cudaEvent_t ev;
        CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        h->chunk_ev.push_back(ev);
    }
    for (int c = 0; c < total; c++) {
        CK(cudaMemcpyAsync(ch[c].pin, ch[c].src, ch[c].len, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaEventRecord(h->chunk_ev[c], h->stream));
    }
    std::atomic<int> next_touch(0), next(0), touch_done(0);
    std::atomic<int> failed(0);
    std::atomic<int> nthreads(0);  // published once every thread that will run has been started
    const int device = h->device;
    auto work = [&]() {
        cudaSetDevice(device);
        while (nthreads.load(std::memory_order_acquire) == 0) std::this_thread::yield();
        for (;;) {  // first touch of the output pages, one write per 4 KB page
            const int c = next_touch.fetch_add(1);
            if (c >= total) break;
            volatile char* p = ch[c].dst;
            for (size_t i = 0; i < ch[c].len; i += 4096) p[i] = 0;
        }
        // no thread may write results before every thread has stopped writing zeros
        touch_done.fetch_add(1, std::memory_order_acq_rel);
        while (touch_done.load(std::memory_order_acquire) < nthreads.load(std::memory_order_acquire))
            std::this_thread::yield();
        for (;;) {
            const int c = next.fetch_add(1);
            if (c >= total) break;
            if (cudaEventSynchronize(h->chunk_ev[c]) != cudaSuccess) {
                failed.store(1);
                break;
            }
            memcpy(ch[c].dst, ch[c].pin, ch[c].len);
        }
    };
    std::vector<std::thread> pool;
    const int started = start_threads(pool, h->host_threads, work);
    nthreads.store(started > 0 ? started : 1, std::memory_order_release);
    if (started == 0) work();  // no thread could be started: copy inline
    for (auto& t : pool) t.join();
    if (failed.load()) return fail(h, UGSM_ERR_CUDA, "staged D2H failed: %s", cudaGetErrorString(cudaGetLastError()));
    return UGSM_OK;
}

// Are these host buffers worth staging?  (all pageable, big enough, copy threads enabled)
bool want_staging(ugsm_handle* h, const HostXfer* x, int n, bool host_is_src) {
    if (const char* e = getenv("UGSM_HOST_THREADS")) h->host_threads = std::min(64, atoi(e));  // tuning override
    if (h->host_threads < 2) return false;
    size_t total = 0;
    for (int i = 0; i < n; i++) {
        total += x[i].bytes;
        if (!is_pageable(host_is_src ? x[i].src : x[i].dst)) return false;
    }
    return total >= 4 * HOST_CHUNK;
}

}  // namespace

// =============================================================================== C-ABI
extern "C" {

const char* ugsm_version(void) { return "ugsm-b200 0.1 (sm_100a)"; }

void ugsm_config_default(ugsm_config* cfg) {
    if (!cfg) return;
    cfg->levels = UGSM_MAX_LEVELS;
    cfg->fovea_level = 7;
    cfg->device = -1;
    cfg->iters_cutoff = 22;
    cfg->smooth_passes = 5;
    cfg->smooth_passes_fine = 10;
    cfg->max_iters = 0;
    cfg->use_graph = 1;
}

int ugsm_query_dims(const ugsm_config* cfg, int rows, int cols, int mode, ugsm_dims* out) {
    if (!cfg || !out) return UGSM_ERR_INVALID;
    return compute_dims(*cfg, rows, cols, mode, out);
}

int ugsm_create(const ugsm_config* cfg, ugsm_handle** out) {
    ugsm_handle* h = nullptr;
    if (!out) return fail(nullptr, UGSM_ERR_INVALID, "out == NULL");
    *out = nullptr;
    ugsm_config c;
    if (cfg)
        c = *cfg;
    else
        ugsm_config_default(&c);
    if (!validate_cfg(c)) return fail(nullptr, UGSM_ERR_INVALID, "invalid config");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev < 1)
        return fail(nullptr, UGSM_ERR_NO_DEVICE, "no CUDA device: %s (this library has no CPU fallback)",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    int dev = c.device;
    if (dev < 0) {
        if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    }
    if (dev >= ndev) return fail(nullptr, UGSM_ERR_INVALID, "device %d out of range (%d devices)", dev, ndev);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, dev)) != cudaSuccess)
        return fail(nullptr, UGSM_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, UGSM_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", dev,
                    prop.major, prop.minor);
    if (const char* e = getenv("UGSM_USE_GRAPH")) c.use_graph = std::min(2, std::max(0, atoi(e)));  // override for profiling runs
    h = new (std::nothrow) ugsm_handle();
    if (!h) return fail(nullptr, UGSM_ERR_NOMEM, "out of host memory");
    h->cfg = c;
    h->device = dev;
    h->sm_count = prop.multiProcessorCount;
    make_taps(h->taps);
    memset(&h->timing, 0, sizeof h->timing);
    if ((e = cudaSetDevice(dev)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking)) != cudaSuccess) {
        delete h;
        return fail(nullptr, UGSM_ERR_CUDA, "stream creation: %s", cudaGetErrorString(e));
    }
    h->stream = h->own_stream;
    if ((e = cudaMalloc(&h->d_flags, (UGSM_MAX_LEVELS + 1) * sizeof(int))) != cudaSuccess ||
        (e = cudaMemset(h->d_flags, 0, (UGSM_MAX_LEVELS + 1) * sizeof(int))) != cudaSuccess) {
        delete h;
        return fail(nullptr, UGSM_ERR_CUDA, "flag allocation: %s", cudaGetErrorString(e));
    }
    for (auto& ev : h->ev) {
        if ((e = cudaEventCreate(&ev)) != cudaSuccess) {
            delete h;
            return fail(nullptr, UGSM_ERR_CUDA, "event creation: %s", cudaGetErrorString(e));
        }
    }
    for (auto& ev : h->kev) {
        if ((e = cudaEventCreate(&ev)) != cudaSuccess) {
            delete h;
            return fail(nullptr, UGSM_ERR_CUDA, "event creation: %s", cudaGetErrorString(e));
        }
    }
    *out = h;
    return UGSM_OK;
}

int ugsm_destroy(ugsm_handle* h) {
    if (!h) return UGSM_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    release_shape(h);
    for (int k = 0; k < 2; k++)
        if (h->d_rgb[k]) cudaFree(h->d_rgb[k]);
    for (int k = 0; k < 3; k++)
        if (h->d_out[k]) cudaFree(h->d_out[k]);
    for (auto& ev : h->ev)
        if (ev) cudaEventDestroy(ev);
    for (auto& ev : h->kev)
        if (ev) cudaEventDestroy(ev);
    for (auto& ev : h->cloud_ev)
        if (ev) cudaEventDestroy(ev);
    if (h->cloud_buf) cudaFree(h->cloud_buf);
    if (h->d_flags) cudaFree(h->d_flags);

    if (h->pin_in) cudaFreeHost(h->pin_in);
    if (h->pin_out) cudaFreeHost(h->pin_out);
    for (auto& ev : h->chunk_ev)
        if (ev) cudaEventDestroy(ev);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
    return UGSM_OK;
}

const char* ugsm_last_error(const ugsm_handle* h) { return h ? h->err.c_str() : g_create_error; }

int ugsm_set_stream(ugsm_handle* h, void* cuda_stream) {
    if (!h) return UGSM_ERR_INVALID;
    h->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->own_stream;
    return UGSM_OK;
}

int ugsm_get_timing(const ugsm_handle* h, ugsm_timing* out) {
    if (!h || !out) return UGSM_ERR_INVALID;
    *out = h->timing;
    return UGSM_OK;
}

int ugsm_match_full_device(ugsm_handle* h, const uint8_t* d_rgbL, const uint8_t* d_rgbR, int rows, int cols,
                           size_t step, float* dH, float* dV, float* dC) {
    if (!h) return UGSM_ERR_INVALID;
    if (!dH || !dV || !dC) return fail(h, UGSM_ERR_INVALID, "null output pointer");
    int rc = run_pair(h, d_rgbL, d_rgbR, rows, cols, step, 0);
    if (rc) return rc;
    rc = emit_full(h, dH, dV, dC);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));
    h->timing.h2d_ms = h->timing.d2h_ms = 0;
    return finish_timing(h);
}

int ugsm_match_foveated_device(ugsm_handle* h, const uint8_t* d_rgbL, const uint8_t* d_rgbR, int rows, int cols,
                               size_t step, float* d_stackH, float* d_stackV, float* d_stackC) {
    if (!h) return UGSM_ERR_INVALID;
    if (!d_stackH || !d_stackV || !d_stackC) return fail(h, UGSM_ERR_INVALID, "null output pointer");
    int rc = run_pair(h, d_rgbL, d_rgbR, rows, cols, step, 1);
    if (rc) return rc;
    rc = emit_stack(h, d_stackH, d_stackV, d_stackC);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));
    h->timing.h2d_ms = h->timing.d2h_ms = 0;
    return finish_timing(h);
}

static int match_host(ugsm_handle* h, const uint8_t* rgbL, const uint8_t* rgbR, int rows, int cols, size_t step,
                      int mode, float* o0, float* o1, float* o2, int* fovW, int* fovH) {
    if (!h) return UGSM_ERR_INVALID;
    if (!rgbL || !rgbR) return fail(h, UGSM_ERR_INVALID, "null pointer argument");
    // an output plane the caller does not want is NULL: it is neither converted nor copied back (all three NULL:
    // the results stay on the device for ugsm_get_level_disparity / ugsm_cloud_from_match)
    const bool want_any = o0 || o1 || o2;
    if (rows < 1 || cols < 1 || step < (size_t)cols * 3) return fail(h, UGSM_ERR_INVALID, "bad rows/cols/step");
    const bool composite = (mode == 2);
    if (composite) mode = 1;
    ugsm_dims d;
    int rc = compute_dims(h->cfg, rows, cols, mode, &d);
    if (rc) return fail(h, rc, "invalid shape/config: rows=%d cols=%d mode=%d", rows, cols, mode);
    CK(cudaSetDevice(h->device));
    // a strided view (cv::Mat ROI, step > cols*3) ends cols*3 bytes into its last row: never touch the rest
    size_t rgb_bytes = step * (size_t)(rows - 1) + (size_t)cols * 3;
    size_t out_floats = (mode == 0 || composite) ? (size_t)rows * cols : (size_t)d.stack_rows * d.fov_width;
    rc = ensure_staging(h, rgb_bytes, want_any ? out_floats : 1);
    if (rc) return rc;
    float* outs[3] = {o0, o1, o2};
    const HostXfer xin[2] = {{rgbL, h->d_rgb[0], rgb_bytes}, {rgbR, h->d_rgb[1], rgb_bytes}};
    HostXfer xout[3];
    int nout = 0;
    for (int k = 0; k < 3; k++)
        if (outs[k]) xout[nout++] = HostXfer{h->d_out[k], outs[k], out_floats * sizeof(float)};
    const bool stage_in = want_staging(h, xin, 2, true), stage_out = nout > 0 && want_staging(h, xout, nout, false);
    if (stage_in || stage_out) {
        rc = ensure_pinned(h, stage_in ? staged_bytes(xin, 2) : 0, stage_out ? staged_bytes(xout, nout) : 0);
        if (rc) return rc;
    }
    CK(cudaEventRecord(h->ev[3], h->stream));
    if (stage_in) {
        rc = staged_h2d(h, xin, 2);
        if (rc) return rc;
    } else {
        CK(cudaMemcpyAsync(h->d_rgb[0], rgbL, rgb_bytes, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(h->d_rgb[1], rgbR, rgb_bytes, cudaMemcpyHostToDevice, h->stream));
    }
    rc = run_pair(h, h->d_rgb[0], h->d_rgb[1], rows, cols, step, mode);
    if (rc) return rc;
    if (composite) {
        rc = run_composite(h);
        if (rc) return rc;
        CK(cudaEventRecord(h->ev[2], h->stream));  // the composite belongs to the pair's device time
        if (want_any) rc = emit_composite(h, h->d_out[0], h->d_out[1], h->d_out[2]);
    } else if (want_any) {
        rc = mode == 0 ? emit_full(h, o0 ? h->d_out[0] : nullptr, o1 ? h->d_out[1] : nullptr, o2 ? h->d_out[2] : nullptr)
                       : emit_stack(h, h->d_out[0], h->d_out[1], h->d_out[2]);
    }
    if (rc) return rc;
    if (stage_out) {
        rc = staged_d2h(h, xout, nout);
        if (rc) return rc;
    } else {
        for (int k = 0; k < 3; k++)
            if (outs[k])
                CK(cudaMemcpyAsync(outs[k], h->d_out[k], out_floats * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    rc = finish_timing(h);
    if (rc) return rc;
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, h->ev[3], h->ev[0]));
    h->timing.h2d_ms = ms;
    h->timing.d2h_ms = 0;
    if (fovW) *fovW = d.fov_width;
    if (fovH) *fovH = d.fov_height;
    return UGSM_OK;
}

int ugsm_match_full(ugsm_handle* h, const uint8_t* rgbL, const uint8_t* rgbR, int rows, int cols, size_t step,
                    float* H, float* V, float* C) {
    return match_host(h, rgbL, rgbR, rows, cols, step, 0, H, V, C, nullptr, nullptr);
}

int ugsm_match_foveated(ugsm_handle* h, const uint8_t* rgbL, const uint8_t* rgbR, int rows, int cols, size_t step,
                        float* stackH, float* stackV, float* stackC, int* fovW, int* fovH) {
    return match_host(h, rgbL, rgbR, rows, cols, step, 1, stackH, stackV, stackC, fovW, fovH);
}

int ugsm_match_composite(ugsm_handle* h, const uint8_t* rgbL, const uint8_t* rgbR, int rows, int cols, size_t step,
                         float* H, float* V, float* C) {
    return match_host(h, rgbL, rgbR, rows, cols, step, 2, H, V, C, nullptr, nullptr);
}

int ugsm_match_composite_device(ugsm_handle* h, const uint8_t* d_rgbL, const uint8_t* d_rgbR, int rows, int cols,
                                size_t step, float* dH, float* dV, float* dC) {
    if (!h) return UGSM_ERR_INVALID;
    if (!dH || !dV || !dC) return fail(h, UGSM_ERR_INVALID, "null output pointer");
    int rc = run_pair(h, d_rgbL, d_rgbR, rows, cols, step, 1);
    if (rc) return rc;
    rc = run_composite(h);
    if (rc) return rc;
    CK(cudaEventRecord(h->ev[2], h->stream));
    rc = emit_composite(h, dH, dV, dC);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));
    h->timing.h2d_ms = h->timing.d2h_ms = 0;
    return finish_timing(h);
}

int ugsm_stage_composite(ugsm_handle* h, const float* stackH, const float* stackV, const float* stackC, int rows,
                         int cols, float* H, float* V, float* C) {
    if (!h) return UGSM_ERR_INVALID;
    if (!stackH || !stackV || !stackC || !H || !V || !C) return fail(h, UGSM_ERR_INVALID, "null pointer argument");
    CK(cudaSetDevice(h->device));
    int rc = prepare_shape(h, rows, cols, 1);
    if (rc) return rc;
    const int fl = h->cfg.fovea_level;
    const float* stack[3] = {stackH, stackV, stackC};
    for (int i = 0; i < fl; i++) {
        LevelBuf& b = h->lv[i];
        size_t off = (size_t)i * b.mw * b.mh;
        for (int k = 0; k < 3; k++) CK(up2d(b.st[k], b.mpitch, stack[k] + off, b.mw, b.mh, h->stream));
    }
    rc = run_composite(h);
    if (rc) return rc;
    h->last_kind = -1;  // staged planes, no image on the device: nothing for ugsm_cloud_from_match
    float* out[3] = {H, V, C};
    for (int k = 0; k < 3; k++) CK(down2d(out[k], h->comp[0][k], h->lv[0].pitch, h->lv[0].w, h->lv[0].h, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return UGSM_OK;
}

// ---- pinned host memory for hosts without a CUDA toolchain of their own ----
int ugsm_host_alloc(size_t bytes, void** out) {
    if (!out || bytes == 0) return UGSM_ERR_INVALID;
    *out = nullptr;
    cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();
        snprintf(g_create_error, sizeof g_create_error, "ugsm_host_alloc: %s", cudaGetErrorString(e));
        *out = nullptr;
        return e == cudaErrorMemoryAllocation ? UGSM_ERR_NOMEM : UGSM_ERR_NO_DEVICE;
    }
    return UGSM_OK;
}

int ugsm_host_free(void* p) {
    if (!p) return UGSM_OK;
    return cudaFreeHost(p) == cudaSuccess ? UGSM_OK : UGSM_ERR_CUDA;
}

// ---- point cloud entry points (K5) ----
int ugsm_cloud_count(int rows, int cols, int sampling, int64_t* npoints) {
    if (rows < 1 || cols < 1 || sampling < 1 || !npoints) return UGSM_ERR_INVALID;
    *npoints = (int64_t)div_up(cols, sampling) * div_up(rows, sampling);
    return UGSM_OK;
}

int ugsm_cloud_full_device(ugsm_handle* h, const ugsm_camera_pair* cam, const float* dH, const float* dV,
                           int pitch_floats, int rows, int cols, const uint8_t* d_img, size_t step, int bgr,
                           int sampling, float* d_points, int point_stride, int64_t* npoints) {
    if (!h) return UGSM_ERR_INVALID;
    if (pitch_floats < cols || step < (size_t)cols * 3) return fail(h, UGSM_ERR_INVALID, "bad pitch/step");
    CK(cudaSetDevice(h->device));
    int rc = launch_cloud(h, cam, dH, dV, pitch_floats, rows, cols, -1, rows, cols, d_img, step, bgr, sampling,
                          d_points, point_stride, npoints);
    if (rc) return rc;
    return finish_cloud(h);
}

// Host planes / image / points: staged through one grow-only device buffer.
static int cloud_host(ugsm_handle* h, const ugsm_camera_pair* cam, const float* X, const float* Y, int rows, int cols,
                      int src_level, int im_rows, int im_cols, const uint8_t* img, size_t step, int bgr, int sampling,
                      float* points, int stride, int64_t* npoints) {
    if (!h) return UGSM_ERR_INVALID;
    if (!cam || !X || !Y || !img || !points) return fail(h, UGSM_ERR_INVALID, "null pointer argument");
    if (rows < 1 || cols < 1 || im_rows < 1 || im_cols < 1 || sampling < 1 || step < (size_t)im_cols * 3)
        return fail(h, UGSM_ERR_INVALID, "bad rows/cols/step/sampling");
    if (stride != 4 && stride != 8) return fail(h, UGSM_ERR_INVALID, "point stride must be 4 or 8 floats");
    CK(cudaSetDevice(h->device));
    const size_t plane = ((size_t)rows * cols * sizeof(float) + 255) / 256 * 256;
    const size_t img_bytes = (step * (size_t)im_rows + 255) / 256 * 256;
    const size_t n = (size_t)div_up(cols, sampling) * div_up(rows, sampling);
    const size_t pts_bytes = n * stride * sizeof(float);
    int rc = ensure_cloud_staging(h, 2 * plane + img_bytes + pts_bytes);
    if (rc) return rc;
    char* base = static_cast<char*>(h->cloud_buf);
    float* dX = reinterpret_cast<float*>(base);
    float* dY = reinterpret_cast<float*>(base + plane);
    uint8_t* d_img = reinterpret_cast<uint8_t*>(base + 2 * plane);
    float* d_pts = reinterpret_cast<float*>(base + 2 * plane + img_bytes);
    const size_t plane_bytes = (size_t)rows * cols * sizeof(float);
    const HostXfer xin[3] = {{X, dX, plane_bytes}, {Y, dY, plane_bytes}, {img, d_img, step * (size_t)(im_rows - 1) + (size_t)im_cols * 3}};
    const HostXfer xout[1] = {{d_pts, points, pts_bytes}};
    const bool stage_in = want_staging(h, xin, 3, true), stage_out = want_staging(h, xout, 1, false);
    if (stage_in || stage_out) {
        rc = ensure_pinned(h, stage_in ? staged_bytes(xin, 3) : 0, stage_out ? staged_bytes(xout, 1) : 0);
        if (rc) return rc;
    }
    if (stage_in) {
        rc = staged_h2d(h, xin, 3);
        if (rc) return rc;
    } else {
        for (int k = 0; k < 3; k++) CK(cudaMemcpyAsync(xin[k].dst, xin[k].src, xin[k].bytes, cudaMemcpyHostToDevice, h->stream));
    }
    rc = launch_cloud(h, cam, dX, dY, cols, rows, cols, src_level, im_rows, im_cols, d_img, step, bgr, sampling, d_pts,
                      stride, npoints);
    if (rc) return rc;
    if (stage_out) {
        rc = staged_d2h(h, xout, 1);
        if (rc) return rc;
    } else {
        CK(cudaMemcpyAsync(points, d_pts, pts_bytes, cudaMemcpyDeviceToHost, h->stream));
    }
    return finish_cloud(h);
}

int ugsm_cloud_full(ugsm_handle* h, const ugsm_camera_pair* cam, const float* H, const float* V, int rows, int cols,
                    const uint8_t* img, size_t step, int bgr, int sampling, float* points, int point_stride,
                    int64_t* npoints) {
    return cloud_host(h, cam, H, V, rows, cols, -1, rows, cols, img, step, bgr, sampling, points, point_stride,
                      npoints);
}

int ugsm_cloud_foveated(ugsm_handle* h, const ugsm_camera_pair* cam, const float* stackH, const float* stackV,
                        int fov_rows, int fov_cols, int src_level, int im_rows, int im_cols, const uint8_t* img,
                        size_t step, int bgr, int sampling, float* points, int point_stride, int64_t* npoints) {
    if (!h) return UGSM_ERR_INVALID;
    if (!stackH || !stackV) return fail(h, UGSM_ERR_INVALID, "null pointer argument");
    if (src_level < 0 || src_level >= h->cfg.fovea_level || fov_rows < 1 || fov_cols < 1)
        return fail(h, UGSM_ERR_INVALID, "src_level %d outside the stack", src_level);
    // the kernel maps fovea coordinates into the image through the level geometry: a fovea of any other size would
    // read the image out of bounds (a sticky device fault for every handle of the process)
    ugsm_dims d;
    if (compute_dims(h->cfg, im_rows, im_cols, 1, &d) != UGSM_OK || fov_cols != d.fov_width || fov_rows != d.fov_height)
        return fail(h, UGSM_ERR_INVALID, "fovea %dx%d does not match the %dx%d image at fovea level %d", fov_cols,
                    fov_rows, im_cols, im_rows, h->cfg.fovea_level);
    const size_t off = (size_t)src_level * fov_rows * fov_cols;  // PC:900: row yy + fovH * srcLevel
    return cloud_host(h, cam, stackH + off, stackV + off, fov_rows, fov_cols, src_level, im_rows, im_cols, img, step,
                      bgr, sampling, points, point_stride, npoints);
}

int ugsm_cloud_from_match(ugsm_handle* h, const ugsm_camera_pair* cam, int src_level, int sampling, float* points,
                          int point_stride, int points_on_device, int64_t* npoints) {
    if (!h) return UGSM_ERR_INVALID;
    if (!cam || !points) return fail(h, UGSM_ERR_INVALID, "null pointer argument");
    if (h->last_kind < 0 || !h->last_rgbL) return fail(h, UGSM_ERR_INVALID, "no match on this handle yet");
    if (point_stride != 4 && point_stride != 8) return fail(h, UGSM_ERR_INVALID, "point stride must be 4 or 8 floats");
    if (sampling < 1) return fail(h, UGSM_ERR_INVALID, "bad sampling");
    CK(cudaSetDevice(h->device));
    const float *dX, *dY;
    int pitch, rows, cols, lvl = -1;
    if (src_level < 0) {
        if (h->last_kind == 0) {  // full pyramid: level-0 state
            dX = h->lv[0].st[0], dY = h->lv[0].st[1], pitch = h->lv[0].mpitch;
        } else if (h->last_kind == 2) {  // hierarchicalDisparity composite
            dX = h->comp[0][0], dY = h->comp[0][1], pitch = h->lv[0].pitch;
        } else {
            return fail(h, UGSM_ERR_INVALID,
                        "the last match was foveated: pass a stack level (or use ugsm_match_composite)");
        }
        rows = h->rows, cols = h->cols;
    } else {
        if (h->last_kind == 0) return fail(h, UGSM_ERR_INVALID, "the last match was not foveated");
        if (src_level >= h->cfg.fovea_level)
            return fail(h, UGSM_ERR_INVALID, "src_level %d outside the stack", src_level);
        const LevelBuf& b = h->lv[src_level];
        dX = b.st[0], dY = b.st[1], pitch = b.mpitch, rows = b.mh, cols = b.mw, lvl = src_level;
    }
    const size_t n = (size_t)div_up(cols, sampling) * div_up(rows, sampling);
    const size_t pts_bytes = n * point_stride * sizeof(float);
    float* d_pts = points;
    if (!points_on_device) {
        int rc = ensure_cloud_staging(h, pts_bytes);
        if (rc) return rc;
        d_pts = static_cast<float*>(h->cloud_buf);
    }
    // the matcher's input images are RGB8 (UG_GPU_matcher.cpp:513-514)
    int rc = launch_cloud(h, cam, dX, dY, pitch, rows, cols, lvl, h->rows, h->cols, h->last_rgbL, h->last_step, 0,
                          sampling, d_pts, point_stride, npoints);
    if (rc) return rc;
    if (!points_on_device) {
        const HostXfer xout[1] = {{d_pts, points, pts_bytes}};
        if (want_staging(h, xout, 1, false)) {
            rc = ensure_pinned(h, 0, staged_bytes(xout, 1));
            if (rc) return rc;
            rc = staged_d2h(h, xout, 1);
            if (rc) return rc;
        } else {
            CK(cudaMemcpyAsync(points, d_pts, pts_bytes, cudaMemcpyDeviceToHost, h->stream));
        }
    }
    return finish_cloud(h);
}

int ugsm_cloud_last_ms(const ugsm_handle* h, float* ms) {
    if (!h || !ms) return UGSM_ERR_INVALID;
    *ms = h->cloud_ms;
    return UGSM_OK;
}

// ---- range maps (getRangePoint / getRangePointFOV, PC:951-1019): z of every pixel, dense row-major ----
static int range_host(ugsm_handle* h, const ugsm_camera_pair* cam, const float* X, const float* Y, int rows, int cols,
                      int src_level, int im_rows, int im_cols, float* range) {
    if (!h) return UGSM_ERR_INVALID;
    if (!cam || !X || !Y || !range) return fail(h, UGSM_ERR_INVALID, "null pointer argument");
    if (rows < 1 || cols < 1 || im_rows < 1 || im_cols < 1) return fail(h, UGSM_ERR_INVALID, "bad rows/cols");
    CK(cudaSetDevice(h->device));
    const size_t plane_bytes = (size_t)rows * cols * sizeof(float), plane = (plane_bytes + 255) / 256 * 256;
    int rc = ensure_cloud_staging(h, 3 * plane);
    if (rc) return rc;
    char* base = static_cast<char*>(h->cloud_buf);
    float* dX = reinterpret_cast<float*>(base);
    float* dY = reinterpret_cast<float*>(base + plane);
    float* dZ = reinterpret_cast<float*>(base + 2 * plane);
    const HostXfer xin[2] = {{X, dX, plane_bytes}, {Y, dY, plane_bytes}};
    const HostXfer xout[1] = {{dZ, range, plane_bytes}};
    const bool stage_in = want_staging(h, xin, 2, true), stage_out = want_staging(h, xout, 1, false);
    if (stage_in || stage_out) {
        rc = ensure_pinned(h, stage_in ? staged_bytes(xin, 2) : 0, stage_out ? staged_bytes(xout, 1) : 0);
        if (rc) return rc;
    }
    if (stage_in) {
        rc = staged_h2d(h, xin, 2);
        if (rc) return rc;
    } else {
        for (int k = 0; k < 2; k++) CK(cudaMemcpyAsync(xin[k].dst, xin[k].src, xin[k].bytes, cudaMemcpyHostToDevice, h->stream));
    }
    rc = launch_cloud(h, cam, dX, dY, cols, rows, cols, src_level, im_rows, im_cols, nullptr, 0, 0, 1, nullptr, 4,
                      nullptr, dZ);
    if (rc) return rc;
    if (stage_out) {
        rc = staged_d2h(h, xout, 1);
        if (rc) return rc;
    } else {
        CK(cudaMemcpyAsync(range, dZ, plane_bytes, cudaMemcpyDeviceToHost, h->stream));
    }
    return finish_cloud(h);
}

int ugsm_range_full(ugsm_handle* h, const ugsm_camera_pair* cam, const float* H, const float* V, int rows, int cols,
                    float* range) {
    return range_host(h, cam, H, V, rows, cols, -1, rows, cols, range);
}

int ugsm_range_foveated(ugsm_handle* h, const ugsm_camera_pair* cam, const float* stackH, const float* stackV,
                        int fov_rows, int fov_cols, int src_level, int im_rows, int im_cols, float* range) {
    if (!h) return UGSM_ERR_INVALID;
    if (!stackH || !stackV) return fail(h, UGSM_ERR_INVALID, "null pointer argument");
    if (src_level < 0 || src_level >= h->cfg.fovea_level || fov_rows < 1 || fov_cols < 1)
        return fail(h, UGSM_ERR_INVALID, "src_level %d outside the stack", src_level);
    const size_t off = (size_t)src_level * fov_rows * fov_cols;  // PC:994: row yy + fovH * srcLevel
    return range_host(h, cam, stackH + off, stackV + off, fov_rows, fov_cols, src_level, im_rows, im_cols, range);
}

int ugsm_range_from_match(ugsm_handle* h, const ugsm_camera_pair* cam, int src_level, float* range, int range_on_device) {
    if (!h) return UGSM_ERR_INVALID;
    if (!cam || !range) return fail(h, UGSM_ERR_INVALID, "null pointer argument");
    if (h->last_kind < 0) return fail(h, UGSM_ERR_INVALID, "no match on this handle yet");
    CK(cudaSetDevice(h->device));
    const float *dX, *dY;
    int pitch, rows, cols, lvl = -1;
    if (src_level < 0) {
        if (h->last_kind == 0) {
            dX = h->lv[0].st[0], dY = h->lv[0].st[1], pitch = h->lv[0].mpitch;
        } else if (h->last_kind == 2) {
            dX = h->comp[0][0], dY = h->comp[0][1], pitch = h->lv[0].pitch;
        } else {
            return fail(h, UGSM_ERR_INVALID, "the last match was foveated: pass a stack level (or use ugsm_match_composite)");
        }
        rows = h->rows, cols = h->cols;
    } else {
        if (h->last_kind == 0) return fail(h, UGSM_ERR_INVALID, "the last match was not foveated");
        if (src_level >= h->cfg.fovea_level) return fail(h, UGSM_ERR_INVALID, "src_level %d outside the stack", src_level);
        const LevelBuf& b = h->lv[src_level];
        dX = b.st[0], dY = b.st[1], pitch = b.mpitch, rows = b.mh, cols = b.mw, lvl = src_level;
    }
    const size_t bytes = (size_t)rows * cols * sizeof(float);
    float* dZ = range;
    if (!range_on_device) {
        int rc = ensure_cloud_staging(h, bytes);
        if (rc) return rc;
        dZ = static_cast<float*>(h->cloud_buf);
    }
    int rc = launch_cloud(h, cam, dX, dY, pitch, rows, cols, lvl, h->rows, h->cols, nullptr, 0, 0, 1, nullptr, 4, nullptr, dZ);
    if (rc) return rc;
    if (!range_on_device) CK(cudaMemcpyAsync(range, dZ, bytes, cudaMemcpyDeviceToHost, h->stream));
    return finish_cloud(h);
}

int ugsm_get_pyramid_level(ugsm_handle* h, int level, float* left3, float* right3) {
    if (!h) return UGSM_ERR_INVALID;
    if (h->lv.empty() || level < 0 || level >= (int)h->lv.size())
        return fail(h, UGSM_ERR_INVALID, "no pyramid level %d (run a match first)", level);
    CK(cudaSetDevice(h->device));
    LevelBuf& b = h->lv[level];
    size_t n = (size_t)b.w * b.h;
    for (int k = 0; k < 3; k++) {
        if (left3) CK(down2d(left3 + k * n, b.L[k], b.pitch, b.w, b.h, h->stream));
        if (right3) CK(down2d(right3 + k * n, b.R[k], b.pitch, b.w, b.h, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    return UGSM_OK;
}

int ugsm_get_fovea_images(ugsm_handle* h, int level, float* left3, float* right3) {
    if (!h) return UGSM_ERR_INVALID;
    if (h->lv.empty() || level < 0 || level >= (int)h->lv.size())
        return fail(h, UGSM_ERR_INVALID, "no level %d (run a match first)", level);
    CK(cudaSetDevice(h->device));
    LevelBuf& b = h->lv[level];
    size_t n = (size_t)b.mw * b.mh;
    for (int k = 0; k < 3; k++) {
        if (left3) CK(down2d(left3 + k * n, b.Lm[k], b.mpitch, b.mw, b.mh, h->stream));
        if (right3) CK(down2d(right3 + k * n, b.Rm[k], b.mpitch, b.mw, b.mh, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    return UGSM_OK;
}

int ugsm_get_level_disparity(ugsm_handle* h, int level, float* H, float* V, float* C) {
    if (!h) return UGSM_ERR_INVALID;
    if (h->lv.empty() || level < 0 || level >= (int)h->lv.size())
        return fail(h, UGSM_ERR_INVALID, "no level %d (run a match first)", level);
    CK(cudaSetDevice(h->device));
    LevelBuf& b = h->lv[level];
    float* outs[3] = {H, V, C};
    for (int k = 0; k < 3; k++)
        if (outs[k]) CK(down2d(outs[k], b.st[k], b.mpitch, b.mw, b.mh, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return UGSM_OK;
}

#include "ugsm_band_api.inc"
#ifndef UGSM_NO_TEST_API
#include "ugsm_test_api.inc"
#endif

}  // extern "C"
