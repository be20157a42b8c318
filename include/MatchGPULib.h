// MatchGPULib.h -- C++ facade with the reference's class name, method names, argument meaning and
// return layouts (reference: src/gpu_matcher/MatchGPULib.h:6-47), forwarding to the C-ABI in ugsm.h.
//
//   * With -DUGSM_WITH_CV_BRIDGE (a ROS build, cv_bridge present) the cv_bridge::CvImagePtr signatures
//     of the reference are compiled in, so UG_GPU_matcher.cpp builds against this header unchanged.
//   * The raw-pointer overloads are the headless path (no ROS/OpenCV): same semantics, RGB8 interleaved
//     with a row step in bytes (what cv_bridge::toCvCopy(..., RGB8) hands the reference,
//     UG_GPU_matcher.cpp:143-144, MatchGPULib.cpp:315-338).
//
// Ownership is the reference's: match() returns a malloc'd array of 3 malloc'd W x H planes [H, V, C];
// the caller frees each plane, then the array (UG_GPU_matcher.cpp:487-489).  matchStack*() return a
// malloc'd [14][3] array of planes; the caller frees the first getFoveateLevel() levels
// (UG_GPU_matcher.cpp:414-418); ugsmFreeStack() frees all of it.
// Differences, all deliberate: no cudaDeviceReset() per pair (MatchGPULib.cpp:400), no global device
// state, errors are reported (lastError(), NULL return) instead of exit().  match(..., fov = 1) runs the
// foveated matcher and the hierarchicalDisparity composite (MatchGPULib.cpp:355-360, 2589-2701).
#ifndef UGSM_MATCHGPULIB_H
#define UGSM_MATCHGPULIB_H

#include <stddef.h>
#include <stdint.h>

#ifdef UGSM_WITH_CV_BRIDGE
#include <cv_bridge/cv_bridge.h>
#endif

struct ugsm_handle;

class MatchGPULib {
   public:
    bool foveatedmatching;
    int foveatelevel;
    int fovH;
    int fovW;
    MatchGPULib(int argc, char** argv);  // argv[2] = fovea level (default 7), -device=N honoured
    ~MatchGPULib();
    MatchGPULib(const MatchGPULib&) = delete;
    MatchGPULib& operator=(const MatchGPULib&) = delete;

    int getFoveaWidth();
    int getFoveaHeight();
    int getFoveateLevel();
    void setFoveaWidth(int rows);
    void setFoveaHeight(int cols);
    void setFoveated(int fov);

#ifdef UGSM_WITH_CV_BRIDGE
    float** match(cv_bridge::CvImagePtr cv_ptrL, cv_bridge::CvImagePtr cv_ptrR, int fov);
    int initStack(cv_bridge::CvImagePtr cv_ptrL, cv_bridge::CvImagePtr cv_ptrR);
    float*** matchStack(cv_bridge::CvImagePtr cv_ptrL, cv_bridge::CvImagePtr cv_ptrR);
    float*** matchStackPyramid(cv_bridge::CvImagePtr cv_ptrL, cv_bridge::CvImagePtr cv_ptrR, float*** leftFoveated,
                               float*** rightFoveated);
#endif
    // headless overloads
    float** match(const uint8_t* rgbL, const uint8_t* rgbR, int rows, int cols, size_t step, int fov);
    int initStack(int rows, int cols);
    float*** matchStack(const uint8_t* rgbL, const uint8_t* rgbR, int rows, int cols, size_t step);
    float*** matchStackPyramid(const uint8_t* rgbL, const uint8_t* rgbR, int rows, int cols, size_t step,
                               float*** leftFoveated, float*** rightFoveated);

    int iDivUp(int a, int b);
    void gaussiankernel(float* k);  // the five normalised taps (MatchGPULib.cpp:735-775)

    const char* lastError() const;
    float lastDeviceMs() const;  // device time of the last pair (CUDA events)

   private:
    ugsm_handle* h_;
    int device_;
    int handle_fovea_level_ = 0;  // the foveatelevel the handle was created for
    char err_[512];
    bool ensure(int levels_unused = 0);
    float*** stackImpl(const uint8_t* rgbL, const uint8_t* rgbR, int rows, int cols, size_t step, float*** lf,
                       float*** rf);
};

// Free everything a matchStack*() call returned (all 14 levels).
void ugsmFreeStack(float*** stack);

#endif
