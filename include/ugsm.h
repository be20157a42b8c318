/*
 * ugsm.h -- C-ABI of the B200-native ug_stereomatcher gpu_matcher hot path.
 *
 * This is the drop-in boundary.  Plain pointers and sizes only;
 * no C++ / torch / OpenCV / ROS types.  Every entry point returns an int status
 * (UGSM_OK == 0) and never throws or aborts across the boundary.
 *
 * What each entry point replaces in the reference (paths relative to the
 * reference checkout, src/gpu_matcher/...):
 *
 *   ugsm_create / ugsm_destroy   MatchGPULib::MatchGPULib(argc, argv)   MatchGPULib.cpp:251-265
 *                                (+ the per-pair cudaDeviceReset         MatchGPULib.cpp:400,528,696)
 *   ugsm_match_full              MatchGPULib::match(L, R, fov=0)        MatchGPULib.cpp:303-403
 *   ugsm_match_foveated          MatchGPULib::matchStackPyramid(...)    MatchGPULib.cpp:534-700
 *                                MatchGPULib::matchStack(...)           MatchGPULib.cpp:429-531
 *   ugsm_query_dims              MatchGPULib::initStack + getFovea*     MatchGPULib.cpp:268-278,406-426
 *   ugsm_stage_*                 the 25 fine-grained extern "C" launch wrappers
 *                                declared at MatchGPULib.cpp:45-247 / MatchLib_common.h:35-74
 *                                (fused here; exposed per fused kernel for parity tests)
 *
 * Plane layout on the host side of this ABI: dense row-major float32, pitch == width,
 * exactly what the reference hands to its ROS node (UG_GPU_matcher.cpp:430-471).
 * RGB inputs: uint8 interleaved R,G,B with a row stride `step` in bytes
 * (cv_bridge rgb8, MatchGPULib.cpp:332-338).
 *
 * Sign convention (getPointCloud.cpp:912-913): the pixel matching left (x, y) is
 * right (x + H, y + V).
 */
#ifndef UGSM_H
#define UGSM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UGSM_MAX_LEVELS 14 /* MAX_LEVEL, MatchLib_common.h:13 */
#define UGSM_TIMED_LEVELS 4

enum {
    UGSM_OK = 0,
    UGSM_ERR_INVALID = 1,  /* bad argument */
    UGSM_ERR_CUDA = 2,     /* CUDA runtime error, see ugsm_last_error */
    UGSM_ERR_NO_DEVICE = 3,
    UGSM_ERR_NOMEM = 4
};

/* Iteration / pyramid schedule.  Defaults (ugsm_config_default) are the reference's
 * compile-time constants and file globals. */
typedef struct ugsm_config {
    int32_t levels;             /* pyramid levels; MAX_LEVEL=14, MatchLib_common.h:13 */
    int32_t fovea_level;        /* foveatelevel; argv[2], default 7, MatchGPULib.cpp:259-264 */
    int32_t device;             /* CUDA device ordinal (-device=N, MatchGPULib.cpp:254); -1 = current */
    int32_t iters_cutoff;       /* levelcutoff=22: iterations on levels i>5, MatchGPULib.cpp:37,1741 */
    int32_t smooth_passes;      /* smoothtime=5, MatchGPULib.cpp:38 */
    int32_t smooth_passes_fine; /* 10 on the two finest levels, MatchGPULib.cpp:2259-2261 */
    int32_t max_iters;          /* 0 = reference schedule; >0 caps iterations per level (tests only) */
    int32_t use_graph;          /* 1 = replay the per-pair launch sequence through a CUDA graph (default); 0 = plain
                                   stream launches with per-level / per-launch timing; 2 = graph for levels >= 1, level
                                   0 launched on the stream with timing events around its K1 / K2 launches
                                   (instrumentation: ugsm_timing.k1_ms[0], k2_ms[0], level_ms[0]) */
} ugsm_config;

typedef struct ugsm_dims {
    int32_t levels;
    int32_t width[UGSM_MAX_LEVELS];  /* natural pyramid dims, w[i+1] = (int)(w[i]/1.41421356) */
    int32_t height[UGSM_MAX_LEVELS];
    int32_t match_width[UGSM_MAX_LEVELS];  /* dims the matcher runs at (fovea-cropped in foveated mode) */
    int32_t match_height[UGSM_MAX_LEVELS];
    int32_t iters[UGSM_MAX_LEVELS];         /* mi per level */
    int32_t smooth[UGSM_MAX_LEVELS];        /* S per level */
    int32_t fov_width, fov_height;          /* fovW, fovH (level fovea_level-1 dims) */
    int32_t stack_rows;                     /* fovea_level * fov_height */
    int64_t pixel_iterations;               /* sum over levels of match_w*match_h*mi */
} ugsm_dims;

/* Device-side timing of the last match call (CUDA events on the handle's stream). */
typedef struct ugsm_timing {
    float total_ms;      /* whole device-resident pair (pyramid .. final planes) */
    float pyramid_ms;    /* unpack + pyramid of both images (0 when replayed from a graph) */
    float levels_ms;     /* coarse-to-fine loop (0 when replayed from a graph) */
    float h2d_ms, d2h_ms;
    int32_t launches;    /* kernels launched by the last call */
    float level_ms[UGSM_MAX_LEVELS]; /* use_graph 0: every level; 2: level 0 only; 1: none (all inside the graph) */
    /* Sum of the K1 (correlate) / K2 (smooth) launch durations per level, from CUDA events
     * recorded around every launch on the handle's stream; measured on the UGSM_TIMED_LEVELS
     * finest levels only (the others stay 0 so that events never pace the tiny kernels). */
    float k1_ms[UGSM_MAX_LEVELS];
    float k2_ms[UGSM_MAX_LEVELS];
} ugsm_timing;

typedef struct ugsm_handle ugsm_handle;

void ugsm_config_default(ugsm_config* cfg);

int ugsm_create(const ugsm_config* cfg, ugsm_handle** out);
int ugsm_destroy(ugsm_handle* h);
/* Last error text of this handle (never NULL).  With h == NULL: last creation error. */
const char* ugsm_last_error(const ugsm_handle* h);

/* Page-locked host memory (portable across devices) for hosts that have no CUDA toolchain of their own: buffers
 * from here make the copies of ugsm_match_* plain asynchronous DMA.  Any other host memory works too (pageable
 * buffers are staged through pinned chunks by copy threads).  Failure text: ugsm_last_error(NULL). */
int ugsm_host_alloc(size_t bytes, void** out);
int ugsm_host_free(void* p);

/* Use an external CUDA stream (cudaStream_t passed as void*); NULL restores the handle's own. */
int ugsm_set_stream(ugsm_handle* h, void* cuda_stream);

/* mode: 0 = full pyramid, 1 = foveated.  Pure host arithmetic. */
int ugsm_query_dims(const ugsm_config* cfg, int rows, int cols, int mode, ugsm_dims* out);

/* Full-pyramid match.  Host buffers (pinned or pageable).  H, V, C: rows*cols floats each.  An output plane that
 * is not wanted may be NULL (here and in ugsm_match_foveated / ugsm_match_composite): it is not copied back -- e.g.
 * H and V only, or nothing at all when the consumer is ugsm_cloud_from_match / ugsm_get_level_disparity, which read
 * the planes the match left in device memory. */
int ugsm_match_full(ugsm_handle* h, const uint8_t* rgbL, const uint8_t* rgbR, int rows, int cols, size_t step,
                    float* H, float* V, float* C);

/* Foveated match.  stackH/V/C: (fovea_level*fovH) x fovW floats, level-major rows
 * (UG_GPU_matcher.cpp:293-303).  fovW/fovH may be NULL. */
int ugsm_match_foveated(ugsm_handle* h, const uint8_t* rgbL, const uint8_t* rgbR, int rows, int cols, size_t step,
                        float* stackH, float* stackV, float* stackC, int* fovW, int* fovH);

/* Same two calls with every buffer already resident in device memory (the throughput
 * path: no host<->device copies inside).  Output planes are dense (pitch == width). */
int ugsm_match_full_device(ugsm_handle* h, const uint8_t* d_rgbL, const uint8_t* d_rgbR, int rows, int cols,
                           size_t step, float* dH, float* dV, float* dC);
int ugsm_match_foveated_device(ugsm_handle* h, const uint8_t* d_rgbL, const uint8_t* d_rgbR, int rows, int cols,
                               size_t step, float* d_stackH, float* d_stackV, float* d_stackC);

/* Foveated match + hierarchical composite = MatchGPULib::match(L, R, fov = 1) (MatchGPULib.cpp:355-360,
 * hierarchicalDisparity MatchGPULib.cpp:2589-2701): full-resolution H, V, C (rows*cols floats each) in which
 * every level's fovea is pasted over the x sqrt2 upsampled composite of the coarser levels. */
int ugsm_match_composite(ugsm_handle* h, const uint8_t* rgbL, const uint8_t* rgbR, int rows, int cols, size_t step,
                         float* H, float* V, float* C);
int ugsm_match_composite_device(ugsm_handle* h, const uint8_t* d_rgbL, const uint8_t* d_rgbR, int rows, int cols,
                                size_t step, float* dH, float* dV, float* dC);

/* After a foveated match: copy the image foveas the reference exports through
 * leftFov/rightFov (MatchGPULib.cpp:644-647) -- level `level`, 3 planes of
 * match_width x match_height floats each, host buffers. */
int ugsm_get_fovea_images(ugsm_handle* h, int level, float* left3, float* right3);

/* After any match: copy pyramid level `level` of the left/right image (3 dense planes each). */
int ugsm_get_pyramid_level(ugsm_handle* h, int level, float* left3, float* right3);
/* After any match: copy the (H,V,C) state the matcher left on level `level`. */
int ugsm_get_level_disparity(ugsm_handle* h, int level, float* H, float* V, float* C);

int ugsm_get_timing(const ugsm_handle* h, ugsm_timing* out);

/* ---- per-kernel stage entry points (host buffers; used by the parity tests and the
 * roofline bench; they run exactly the kernels the match calls run) ---------------- */

/* K0: RGB8 -> planar f32 + Gaussian pyramid.  out3[l] for l in [0,levels): 3 dense planes
 * of width[l]*height[l]; entries may be NULL to skip the copy-out. */
int ugsm_stage_pyramid(ugsm_handle* h, const uint8_t* rgb, int rows, int cols, size_t step, float* const* out3);

/* K1: one fused warp + 5-offset NCC^2 + RGB mean + parabola fit + confidence blend.
 * L3/R3: 3 dense planes each.  first_coarsest != 0 skips the confidence blend
 * (MatchGPULib.cpp:2223).  Optional debug outputs (may be NULL): warp source indices
 * and the PolyDisparity branch taken (0 flat, 1 concave, 2 concave-rescaled). */
int ugsm_stage_correlate(ugsm_handle* h, const float* L3, const float* R3, const float* H, const float* V,
                         const float* C, int width, int height, float threshold, int first_coarsest, float* Hn,
                         float* Vn, float* Cn, int32_t* dbg_srcx, int32_t* dbg_srcy, uint8_t* dbg_branch_h,
                         uint8_t* dbg_branch_v);

/* K2: `passes` confidence-weighted cross smoothing passes followed by the 3x3 box filter, in place. */
int ugsm_stage_smooth(ugsm_handle* h, float* H, float* V, float* C, int width, int height, int passes);

/* K3: nearest-neighbour x sqrt(2) upsample of (H,V,C) to the next finer level; crop_x/crop_y/out_w/out_h
 * select the centre crop of the foveated variant (MatchGPULib.cpp:1595-1655); for the plain variant
 * pass crop 0,0 and out dims == fine dims. */
int ugsm_stage_upsample(ugsm_handle* h, const float* H, const float* V, const float* C, int width, int height,
                        int fine_w, int fine_h, int crop_x, int crop_y, int out_w, int out_h, float* Ho, float* Vo,
                        float* Co);

/* The parabola fit K1 uses (PolyDisparity, MatchLib.cu:790-843) on n (l, c, r) triples. */
int ugsm_stage_poly(ugsm_handle* h, const float* l, const float* c, const float* r, int n, float T, float* offset,
                    float* conf, uint8_t* branch);
/* K4 chain: hierarchicalDisparity on a given foveated stack (layout of ugsm_match_foveated). */
int ugsm_stage_composite(ugsm_handle* h, const float* stackH, const float* stackV, const float* stackC, int rows,
                         int cols, float* H, float* V, float* C);

/* One whole level: `iters` iterations of K1+K2 with the reference threshold schedule. */
int ugsm_stage_match_level(ugsm_handle* h, const float* L3, const float* R3, float* H, float* V, float* C, int width,
                           int height, int iters, int passes, int coarsest);

/* Micro-benchmark of K1 / K2 on device-resident synthetic planes: runs `reps` launches and
 * returns the average per-launch time from CUDA events on the handle's stream. */
int ugsm_bench_correlate(ugsm_handle* h, int width, int height, int reps, float* avg_ms);
int ugsm_bench_smooth(ugsm_handle* h, int width, int height, int passes, int reps, float* avg_ms);
/* Launch floor of the stream: average time of `reps` back-to-back empty kernel launches. */
/* K0 timing on synthetic planes: kind 0 = left-energy (dense, 3 planes), 1 = blur + odd-sample decimation
 * (6 planes), 2 = blur + sqrt2 decimation (6 planes); average launch time over `reps`. */
int ugsm_bench_pyramid(ugsm_handle* h, int width, int height, int kind, int reps, float* avg_ms);
int ugsm_bench_null(ugsm_handle* h, int reps, float* avg_ms);

/* ---- row-band execution of the finest levels across GPUs ------------------------
 * One process per GPU, every rank holds the whole pyramid (levels >= split_levels are computed
 * replicated); for the levels below split_levels each rank runs K1/K2 only on its band of rows and
 * the caller exchanges halo rows of (H,V,C) between iterations (NVLink P2P / NCCL; see
 * ug_stereomatcher_b200/rowband.py).  All calls enqueue on the handle's stream and do not
 * synchronise; buffers are device pointers. */
typedef struct ugsm_band_level {
    int32_t width, height, pitch; /* level dims, pitch in floats */
    int32_t iters, smooth;        /* iterations and smoothing passes of this level */
    float* state[3];              /* H, V, C planes of the level (pitched), the ones halos are exchanged on */
} ugsm_band_level;

/* Pyramid of both images + levels [levels-1 .. split_levels] + upsample into level split_levels-1. */
int ugsm_band_begin(ugsm_handle* h, const uint8_t* d_rgbL, const uint8_t* d_rgbR, int rows, int cols, size_t step,
                    int split_levels);
int ugsm_band_level_info(ugsm_handle* h, int level, ugsm_band_level* out);
/* Iteration m (1-based) of `level` for rows [y_begin, y_end): K1 on the band extended by `ext` rows
 * (clipped to the image; ext >= smooth+1 so that K2 needs no exchange of K1's output), then K2. */
int ugsm_band_iteration(ugsm_handle* h, int level, int m, int y_begin, int y_end, int ext);
/* Upsample rows [y_begin, y_end) of level-1 from `level` (K3). */
int ugsm_band_upsample(ugsm_handle* h, int level, int y_begin, int y_end);
/* Halo staging: copy rows [y_begin, y_end) of the level's H, V, C planes into / out of one contiguous
 * device buffer of 3 * (y_end - y_begin) * pitch floats (plane-major). */
int ugsm_band_pack(ugsm_handle* h, int level, int y_begin, int y_end, float* d_buf);
int ugsm_band_unpack(ugsm_handle* h, int level, int y_begin, int y_end, const float* d_buf);
/* Copy rows [y_begin, y_end) of the level-0 result into dense (pitch == width) output planes. */
int ugsm_band_emit(ugsm_handle* h, float* dH, float* dV, float* dC, int y_begin, int y_end);

/* The same schedule driven from C++ inside the library (csrc/ugsm_rowband.cu): one pair split into row bands over
 * n handles, one per GPU of the box (devices[r] = CUDA ordinal of handles[r]; several ranks on one device emulate
 * the split on a single GPU).  One host thread per rank; halo rows travel straight from one GPU's level planes into
 * its neighbour's (cudaMemcpy2DAsync over NVLink peer access, ordered by events), every rank writes its band of the
 * result into the dense planes dH, dV, dC on devices[0].  d_rgbL[r] / d_rgbR[r]: the image pair resident on
 * devices[r].  device_ms (may be NULL): time on rank 0's stream from the first launch to the assembled result.
 * Bit-identical to ugsm_match_full_device.  Error text: ugsm_rowband_last_error(). */
int ugsm_rowband_match(ugsm_handle* const* handles, const int* devices, int n, const uint8_t* const* d_rgbL,
                       const uint8_t* const* d_rgbR, int rows, int cols, size_t step, int split_levels, float* dH,
                       float* dV, float* dC, float* device_ms);
const char* ugsm_rowband_last_error(void);

const char* ugsm_version(void);

/* ---- Point cloud from the disparity planes.  K5 replaces the per-pixel host loops of
 * the reference's point-cloud node, src/pointcloud/getPointCloud.cpp (PC): get3DPoint PC:886-945,
 * doReconstructionRGB PC:741-782, doReconstructionRGB_FOV PC:697-740 with mapX/Ycoord PC:375-421 and the
 * margins PC:432-484.  Results are bit-identical to that code (x86-64 host build, no FMA) for finite
 * disparities; the "resized" clouds (PC:784-884, OpenCV bicubic resize) are not provided. ---- */
typedef struct ugsm_camera_pair {
    double P1[12]; /* sensor_msgs/CameraInfo.P of the left camera, row-major 3x4 (PC:1158) */
    double P2[12]; /* ... of the right camera (PC:1159) */
} ugsm_camera_pair;

/* Number of points a rows x cols plane yields at this sampling: ceil(cols/s) * ceil(rows/s) (PC:746-759). */
int ugsm_cloud_count(int rows, int cols, int sampling, int64_t* npoints);

/* Full-resolution cloud (PC:741-782).  H, V: rows*cols floats (the matcher's output planes); img: the left image,
 * 3 interleaved bytes per pixel, `step` bytes per row, bgr = 1 for B,G,R byte order (the node's cv image), 0 for
 * R,G,B.  points: npoints records of point_stride floats, in the reference's column-major order (record
 * (x/s) * ceil(rows/s) + y/s).  point_stride 4: {x, y, z, rgb}; 8: the pcl::PointXYZRGB memory layout
 * {x, y, z, 1.0f, rgb, 0, 0, 0}.  rgb = the bits of r<<16 | g<<8 | b (PC:770-775).  Host buffers. */
int ugsm_cloud_full(ugsm_handle* h, const ugsm_camera_pair* cam, const float* H, const float* V, int rows, int cols,
                    const uint8_t* img, size_t step, int bgr, int sampling, float* points, int point_stride,
                    int64_t* npoints);
/* Same with every buffer already in device memory; dH/dV rows are pitch_floats apart; d_points 16-byte aligned. */
int ugsm_cloud_full_device(ugsm_handle* h, const ugsm_camera_pair* cam, const float* dH, const float* dV,
                           int pitch_floats, int rows, int cols, const uint8_t* d_img, size_t step, int bgr,
                           int sampling, float* d_points, int point_stride, int64_t* npoints);
/* Foveated cloud of stack level src_level (0 = finest fovea; PC:697-740).  stackH/V: the (levels*fov_rows) x
 * fov_cols stacks ugsm_match_foveated returns; img: the full-resolution im_rows x im_cols left image. */
int ugsm_cloud_foveated(ugsm_handle* h, const ugsm_camera_pair* cam, const float* stackH, const float* stackV,
                        int fov_rows, int fov_cols, int src_level, int im_rows, int im_cols, const uint8_t* img,
                        size_t step, int bgr, int sampling, float* points, int point_stride, int64_t* npoints);
/* Cloud of the pair last matched on this handle, straight from the planes and the left image still resident in
 * device memory (no D2H / H2D of the 16 MP planes).  src_level < 0: the full-resolution planes of the last
 * ugsm_match_full / ugsm_match_composite; src_level >= 0: that level of the last foveated match.  After a
 * *_device match the caller's left image must still be valid.  points: host (points_on_device = 0) or device. */
int ugsm_cloud_from_match(ugsm_handle* h, const ugsm_camera_pair* cam, int src_level, int sampling, float* points,
                          int point_stride, int points_on_device, int64_t* npoints);
/* Device time of the last cloud kernel (CUDA events on the handle's stream). */
int ugsm_cloud_last_ms(const ugsm_handle* h, float* ms);

/* Range maps: what CdynamicCalibration::getRangePoint / getRangePointFOV (PC:951-1019) return at every pixel, i.e. the
 * loops of doReconstruction_resized / doReconstructionFOV_resized (PC:730-751, 806-834) up to, not including, their
 * cv::resize: z of the triangulated point (the two functions share get3DPoint's ZUp / divisor expressions), dense
 * row-major rows x cols (fov_rows x fov_cols) floats as range_map.at<float>(jj, ii).  Bit-identical to that code. */
int ugsm_range_full(ugsm_handle* h, const ugsm_camera_pair* cam, const float* H, const float* V, int rows, int cols,
                    float* range);
int ugsm_range_foveated(ugsm_handle* h, const ugsm_camera_pair* cam, const float* stackH, const float* stackV,
                        int fov_rows, int fov_cols, int src_level, int im_rows, int im_cols, float* range);
/* Same on the planes the last match left on the device (src_level < 0: full / composite; else that stack level);
 * range: host buffer, or device buffer when range_on_device != 0. */
int ugsm_range_from_match(ugsm_handle* h, const ugsm_camera_pair* cam, int src_level, float* range, int range_on_device);

/* ---- Real-image feeder: image list -> TIFF decode -> ring of page-locked pair buffers.
 * Replaces, for batch runs, the reference's camera simulator: Settings::readStringList / nextImage
 * (include/ug_stereomatcher/settings.h:15-68: OpenCV FileStorage XML whose first top-level node is a sequence of
 * paths, read two at a time as left, right; past the end the reader steps back to the next-to-last entry, i.e. it
 * repeats the last pair) and RHcam_node::captureImage (src/utils/publish_images.cpp:128-211: cv::imread(path, 1),
 * i.e. 8-bit, 3 channels, B,G,R byte order, handed on unswapped).  Decoded bytes are identical to cv::imread's for
 * the supported TIFF variants (csrc/ugsm_feeder.cpp).  Relative paths are resolved against the list's directory.
 * Host-only code. ---- */
#define UGSM_FEEDER_END 100 /* ugsm_feeder_next: every pair has been handed out */
typedef struct ugsm_feeder ugsm_feeder;
typedef struct ugsm_feeder_pair {
    const uint8_t* left;  /* rows x cols x 3 bytes (B,G,R), `step` bytes per row, page-locked when info.pinned */
    const uint8_t* right;
    int32_t rows, cols;
    size_t step;
    int64_t index; /* pair number, 0-based, in list order */
    int32_t slot;  /* give back with ugsm_feeder_release once the match call that reads it has returned */
} ugsm_feeder_pair;
typedef struct ugsm_feeder_info_t {
    int32_t rows, cols;
    size_t step;
    int64_t list_images; /* entries in the list */
    int64_t pairs;       /* pairs this feeder delivers */
    int32_t slots;
    int32_t pinned;      /* 1: slots come from ugsm_host_alloc; 0: no CUDA device, ordinary memory */
    double decode_ms;    /* decode time summed over the pairs decoded so far (all threads) */
} ugsm_feeder_info_t;
/* pairs <= 0: the list's own length / 2; pairs beyond the list repeat its last pair (settings.h:34-38).
 * slots: pairs held in the ring; threads: decode threads running ahead of ugsm_feeder_next. */
int ugsm_feeder_open(const char* list_xml, long pairs, int slots, int threads, ugsm_feeder** out);
int ugsm_feeder_info(const ugsm_feeder* f, ugsm_feeder_info_t* out);
/* Blocks until the next pair (list order) is decoded.  Thread-safe: concurrent callers get distinct pairs.
 * UGSM_FEEDER_END after the last pair; a decode error is returned for that pair only. */
int ugsm_feeder_next(ugsm_feeder* f, ugsm_feeder_pair* out);
int ugsm_feeder_release(ugsm_feeder* f, int slot);
int ugsm_feeder_close(ugsm_feeder* f);
/* Error text of the feeder (f == NULL: of the calling thread's last failed open / tiff call). */
const char* ugsm_feeder_last_error(const ugsm_feeder* f);
/* One file: header fields, and the pixels as cv::imread(path, 1) returns them. */
int ugsm_tiff_info(const char* path, int* rows, int* cols, int* samples, int* bits);
int ugsm_tiff_read_bgr8(const char* path, uint8_t* dst, size_t step, int rows, int cols);

#ifdef __cplusplus
}
#endif
#endif /* UGSM_H */
