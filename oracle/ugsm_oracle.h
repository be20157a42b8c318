/*
 * ugsm_oracle.h -- CPU oracle for the ug_stereomatcher gpu_matcher hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (ug_stereomatcher_b200/, the
 * C-ABI library) links, imports or executes this.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * It is a scalar restatement (C99 + OpenMP over rows) of the reference kernels in
 * src/gpu_matcher/MatchLib.cu orchestrated as src/gpu_matcher/MatchGPULib.cpp does;
 * every function cites the reference lines it follows.  Arithmetic policy: each
 * reference kernel stores f32 planes, so every intermediate is rounded to f32; the
 * reference is built with nvcc defaults (CMakeLists.txt:86, -fmad=true), so the
 * multiply-add chains nvcc fuses (checked in the SASS of the reference kernels built
 * for sm_100a) are written as explicit fmaf()/fma() here and nothing
 * else may be contracted (build with -ffp-contract=off).
 *
 * Pinning status: PINNED.  The reference ships no golden vectors or tests (SURVEY.md
 * section 4), so the pins were made here: (a) the reference's own CPU convolution
 * (convolutionSeparable_gold.cpp, compiled in place into oracle/_ref/) and (b) outputs of
 * the UNMODIFIED reference CUDA sources, built for sm_100a through the compat shims in
 * oracle/refbuild/ and run on a B200: MatchGPULib::matchlevel (2- and 22-iteration
 * schedules, coarsest-level variant), convolutionGPU, subsampledimageGPU and subsampleDisp
 * at tile-aligned sizes -- this oracle reproduces every one of them BIT FOR BIT
 * (tests/golden/refcuda_stages.npz, tests/test_refcuda_golden.py).  The reference's full
 * 14-level run is not a usable bit-oracle (out-of-bounds writes at partial tiles), see
 * DESIGN.md section 3.
 */
#ifndef UGSM_ORACLE_H
#define UGSM_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UGSMO_MAX_LEVELS 14

typedef struct ugsmo_config {
    int32_t levels;
    int32_t fovea_level;
    int32_t iters_cutoff;
    int32_t smooth_passes;
    int32_t smooth_passes_fine;
    int32_t max_iters;
} ugsmo_config;

void ugsmo_config_default(ugsmo_config* cfg);
void ugsmo_set_num_threads(int n);
int ugsmo_num_threads(void);

/* MatchGPULib.cpp:735-775 / 344-348 */
void ugsmo_taps(float gauss[5], float box[5]);
/* MatchGPULib.cpp:1224-1228 */
void ugsmo_level_dims(int cols, int rows, int levels, int* w, int* h);
/* MatchGPULib.cpp:1741 / 2257-2261 */
int ugsmo_level_iters(const ugsmo_config* cfg, int i);
int ugsmo_level_smooth(const ugsmo_config* cfg, int i);
/* MatchGPULib.cpp:2299-2306 */
float ugsmo_threshold_after(int mi, int m, float T);

/* MatchGPULib.cpp:332-338 */
void ugsmo_unpack_rgb8(const uint8_t* rgb, int rows, int cols, size_t step, float* planes3);

/* MatchLib.cu:71-156,195-278 with the in-image semantics of convolutionSeparable_gold.cpp:20-75 */
void ugsmo_conv_rows_zero(const float* src, float* dst, int w, int h, const float taps[5]);
void ugsmo_conv_cols_zero(const float* src, float* dst, int w, int h, const float taps[5]);
/* MatchLib.cu:1461-1565 (Gaussian taps) and 1593-1697 (box taps): clamp border */
void ugsmo_conv_rows_clamp(const float* src, float* dst, int w, int h, const float taps[5]);
void ugsmo_conv_cols_clamp(const float* src, float* dst, int w, int h, const float taps[5]);
/* MatchLib.cu:311-339 */
void ugsmo_decimate(const float* src, int w, int h, float scale, float* dst, int w2, int h2);
/* MatchLib.cu:372-401 */
void ugsmo_upsample_disp(const float* src, int w, int h, float* dst, int w2, int h2);
/* MatchLib.cu:499-520 */
void ugsmo_warp(const float* R, const float* H, const float* V, int w, int h, float* dst, int32_t* srcx,
                int32_t* srcy);
/* MatchLib.cu:790-843 */
void ugsmo_poly(float l, float c, float r, float T, float* offset, float* conf, uint8_t* branch);
/* ugsmo_poly on n triples */
void ugsmo_poly_n(const float* l, const float* c, const float* r, int n, float T, float* offset, float* conf,
                  uint8_t* branch);
/* MatchLib.cu:1092-1145 */
void ugsmo_smooth_pass(const float* src, const float* conf, float* dst, int w, int h);

/* MatchGPULib.cpp:1033-1125: planes of one channel; out[l] has dims from ugsmo_level_dims */
void ugsmo_pyramid_plane(const float* plane0, int cols, int rows, int levels, float* const* out);

/* MatchGPULib.cpp:1743-2250 (one iteration up to and including the confidence blend) */
void ugsmo_correlate(const float* L3, const float* R3, const float* H, const float* V, const float* C, int w, int h,
                     float T, int first_coarsest, float* Hn, float* Vn, float* Cn, int32_t* srcx, int32_t* srcy,
                     uint8_t* branch_h, uint8_t* branch_v);
/* MatchGPULib.cpp:2257-2293 + 2361-2412 */
void ugsmo_smooth(float* H, float* V, float* C, int w, int h, int passes);
/* MatchGPULib.cpp:1662-2489 */
void ugsmo_match_level(const float* L3, const float* R3, float* H, float* V, float* C, int w, int h, int iters,
                       int passes, int coarsest);
/* MatchGPULib.cpp:1526-1590 and 1595-1655 (crop variant) */
void ugsmo_upsample_state(const float* H, const float* V, const float* C, int w, int h, int fine_w, int fine_h,
                          int crop_x, int crop_y, int out_w, int out_h, float* Ho, float* Vo, float* Co);

/* MatchGPULib.cpp:303-403 (fov == 0) */
int ugsmo_match_full(const ugsmo_config* cfg, const uint8_t* rgbL, const uint8_t* rgbR, int rows, int cols,
                     size_t step, float* H, float* V, float* C);
/* MatchGPULib.cpp:534-700 + UG_GPU_matcher.cpp:293-303 stack layout */
int ugsmo_match_foveated(const ugsmo_config* cfg, const uint8_t* rgbL, const uint8_t* rgbR, int rows, int cols,
                         size_t step, float* stackH, float* stackV, float* stackC, int* fovW, int* fovH);

/* MatchLib.cu:435-464 (partsubsampleDispKernel): dst = s * src[floor((y+.5)/s)][floor((x+.5)/s)], all float */
void ugsmo_partsubsample_disp(const float* src, int w, int h, float scale, float* dst, int w2, int h2);
/* MatchGPULib.cpp:2589-2701 (hierarchicalDisparity): stack3 = foveated stack (levels 0..fl-1, each 3 planes of
 * fovW x fovH, level-major as ugsmo_match_foveated returns it) -> 3 full-resolution planes */
void ugsmo_composite(const float* stackH, const float* stackV, const float* stackC, int cols, int rows, int levels,
                     int fovea_level, float* H, float* V, float* C);
/* MatchGPULib.cpp:303-403 with fov == 1 */
int ugsmo_match_composite(const ugsmo_config* cfg, const uint8_t* rgbL, const uint8_t* rgbR, int rows, int cols,
                          size_t step, float* H, float* V, float* C);

/* ---- point cloud from H, V (oracle/ugsm_oracle_cloud.c) ---- */
/* getPointCloud.cpp:886-945 (get3DPoint); P1, P2 = CameraInfo.P of the left / right camera, row-major 3x4 */
void ugsmo_triangulate(const double P1[12], const double P2[12], float x1, float y1, float x2, float y2,
                       float out[3]);
int64_t ugsmo_cloud_count(int rows, int cols, int sampling);
/* getPointCloud.cpp:741-782 (doReconstructionRGB); points: count x stride floats, stride 4 = {x,y,z,rgb},
 * stride 8 = pcl::PointXYZRGB layout {x,y,z,1,rgb,0,0,0}; column-major point order; returns the count */
int64_t ugsmo_cloud_full(const double P1[12], const double P2[12], const float* dX, const float* dY, int rows,
                         int cols, const uint8_t* img, size_t step, int bgr, int sampling, float* points,
                         int stride);
/* getPointCloud.cpp:432-484 (destLevel 0) and 375-421 */
void ugsmo_cloud_margins(int im_rows, int im_cols, int fovea_level, int src_level, int* left, int* upper);
float ugsmo_cloud_scale(int src_level);
/* getPointCloud.cpp:697-740 (doReconstructionRGB_FOV) */
int64_t ugsmo_cloud_foveated(const double P1[12], const double P2[12], const float* stackX, const float* stackY,
                             int fov_h, int fov_w, int fovea_level, int src_level, int im_rows, int im_cols,
                             const uint8_t* img, size_t step, int bgr, int sampling, float* points, int stride);

#ifdef __cplusplus
}
#endif
/* Range maps (PC:951-1019): z of the triangulated point at every pixel, row-major. */
int ugsmo_range_full(const double P1[12], const double P2[12], const float* dX, const float* dY, int rows, int cols,
                     float* range);
int ugsmo_range_foveated(const double P1[12], const double P2[12], const float* stackX, const float* stackY, int fov_h,
                         int fov_w, int fovea_level, int src_level, int im_rows, int im_cols, float* range);

#endif
