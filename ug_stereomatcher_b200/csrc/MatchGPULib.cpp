// MatchGPULib.cpp -- the reference-named C++ facade over the C-ABI (see include/MatchGPULib.h).
// Reference being mirrored: src/gpu_matcher/MatchGPULib.cpp:251-700.
#include "../../include/MatchGPULib.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/ugsm.h"

MatchGPULib::MatchGPULib(int argc, char** argv) : h_(nullptr), device_(-1) {
    err_[0] = 0;
    foveatedmatching = 0;  // MatchGPULib.cpp:255-257
    fovH = 0;
    fovW = 0;
    foveatelevel = (argc > 2) ? atoi(argv[2]) : 7;  // MatchGPULib.cpp:259-264
    for (int i = 1; i < argc; i++)                   // findCudaDevice(argc, argv): -device=N (MatchGPULib.cpp:254)
        if (argv[i] && !strncmp(argv[i], "-device=", 8)) device_ = atoi(argv[i] + 8);
}

MatchGPULib::~MatchGPULib() {
    if (h_) ugsm_destroy(h_);
}

int MatchGPULib::getFoveaWidth() { return fovW; }
int MatchGPULib::getFoveaHeight() { return fovH; }
int MatchGPULib::getFoveateLevel() { return foveatelevel; }
int MatchGPULib::iDivUp(int a, int b) { return (a % b != 0) ? (a / b + 1) : (a / b); }
void MatchGPULib::setFoveaWidth(int rows) { fovW = rows; }
void MatchGPULib::setFoveaHeight(int cols) { fovH = cols; }
void MatchGPULib::setFoveated(int fov) { foveatedmatching = fov; }
const char* MatchGPULib::lastError() const { return err_; }

float MatchGPULib::lastDeviceMs() const {
    ugsm_timing t;
    if (!h_ || ugsm_get_timing(h_, &t) != UGSM_OK) return 0.f;
    return t.total_ms;
}

void MatchGPULib::gaussiankernel(float* k) {  // MatchGPULib.cpp:761-775
    k[0] = 0.0816475;
    k[1] = 0.218507;
    k[2] = 0.303281;
    k[3] = 0.218507;
    k[4] = 0.0816475;
    float kern = 0;
    for (int i = 0; i < 5; i++) kern = kern + k[i];
    for (int i = 0; i < 5; i++) k[i] = k[i] / kern;
}

bool MatchGPULib::ensure(int) {
    // foveatelevel is a public field of the reference class: a handle created for another value is replaced
    if (h_ && handle_fovea_level_ != foveatelevel) {
        ugsm_destroy(h_);
        h_ = nullptr;
    }
    if (h_) return true;
    handle_fovea_level_ = foveatelevel;
    ugsm_config cfg;
    ugsm_config_default(&cfg);
    cfg.fovea_level = foveatelevel;
    cfg.device = device_;
    if (ugsm_create(&cfg, &h_) != UGSM_OK) {
        snprintf(err_, sizeof err_, "%s", ugsm_last_error(nullptr));
        fprintf(stderr, "MatchGPULib: %s\n", err_);
        h_ = nullptr;
        return false;
    }
    return true;
}

// MatchGPULib.cpp:406-426: fovea = dims of pyramid level foveatelevel-1.
int MatchGPULib::initStack(int rows, int cols) {
    ugsm_config cfg;
    ugsm_config_default(&cfg);
    cfg.fovea_level = foveatelevel;
    ugsm_dims d;
    if (ugsm_query_dims(&cfg, rows, cols, 1, &d) != UGSM_OK) {
        snprintf(err_, sizeof err_, "initStack: invalid image size %dx%d for fovea level %d", cols, rows, foveatelevel);
        return 1;
    }
    setFoveaHeight(d.fov_height);
    setFoveaWidth(d.fov_width);
    return 0;
}

// MatchGPULib.cpp:303-403
float** MatchGPULib::match(const uint8_t* rgbL, const uint8_t* rgbR, int rows, int cols, size_t step, int fov) {
    foveatedmatching = fov;
    if (!ensure()) return nullptr;
    size_t n = (size_t)rows * cols;
    float** out = (float**)malloc(3 * sizeof(float*));
    for (int k = 0; k < 3; k++) out[k] = (float*)malloc(n * sizeof(float));
    // fov == 1: foveated matching + hierarchicalDisparity composite (MatchGPULib.cpp:355-360)
    int rc = fov ? ugsm_match_composite(h_, rgbL, rgbR, rows, cols, step, out[0], out[1], out[2])
                 : ugsm_match_full(h_, rgbL, rgbR, rows, cols, step, out[0], out[1], out[2]);
    if (rc != UGSM_OK) {
        snprintf(err_, sizeof err_, "%s", ugsm_last_error(h_));
        fprintf(stderr, "MatchGPULib: %s\n", err_);
        for (int k = 0; k < 3; k++) free(out[k]);
        free(out);
        return nullptr;
    }
    return out;
}

// MatchGPULib.cpp:429-700: disp[MAX_LEVEL][3]; levels below the fovea level hold fovW x fovH planes, the
// others their natural dims; leftFov/rightFov receive the image foveas (MatchGPULib.cpp:644-647).
float*** MatchGPULib::stackImpl(const uint8_t* rgbL, const uint8_t* rgbR, int rows, int cols, size_t step, float*** lf,
                                float*** rf) {
    foveatedmatching = 1;
    if (!ensure()) return nullptr;
    ugsm_config cfg;
    ugsm_config_default(&cfg);
    cfg.fovea_level = foveatelevel;
    ugsm_dims d;
    if (ugsm_query_dims(&cfg, rows, cols, 1, &d) != UGSM_OK) {
        snprintf(err_, sizeof err_, "matchStack: invalid image size %dx%d for fovea level %d", cols, rows, foveatelevel);
        return nullptr;
    }
    fovW = d.fov_width;  // MatchGPULib.cpp:1233-1234
    fovH = d.fov_height;
    // no stack planes requested: the per-level planes are fetched below at their own dims, once
    int rc = ugsm_match_foveated(h_, rgbL, rgbR, rows, cols, step, nullptr, nullptr, nullptr, nullptr, nullptr);
    if (rc != UGSM_OK) {
        snprintf(err_, sizeof err_, "%s", ugsm_last_error(h_));
        fprintf(stderr, "MatchGPULib: %s\n", err_);
        return nullptr;
    }
    float*** disp = (float***)malloc(UGSM_MAX_LEVELS * sizeof(float**));
    for (int i = 0; i < UGSM_MAX_LEVELS; i++) {
        size_t n = (size_t)d.match_width[i] * d.match_height[i];
        disp[i] = (float**)malloc(3 * sizeof(float*));
        for (int k = 0; k < 3; k++) disp[i][k] = (float*)malloc(n * sizeof(float));
        ugsm_get_level_disparity(h_, i, disp[i][0], disp[i][1], disp[i][2]);
        if (lf && rf) {
            float* l3 = (float*)malloc(3 * n * sizeof(float));
            float* r3 = (float*)malloc(3 * n * sizeof(float));
            ugsm_get_fovea_images(h_, i, l3, r3);
            float** lp = (float**)malloc(3 * sizeof(float*));
            float** rp = (float**)malloc(3 * sizeof(float*));
            for (int k = 0; k < 3; k++) {
                lp[k] = (float*)malloc(n * sizeof(float));
                rp[k] = (float*)malloc(n * sizeof(float));
                memcpy(lp[k], l3 + k * n, n * sizeof(float));
                memcpy(rp[k], r3 + k * n, n * sizeof(float));
            }
            free(l3);
            free(r3);
            lf[i] = lp;  // the reference replaces the caller's row pointers the same way
            rf[i] = rp;
        }
    }
    return disp;
}

float*** MatchGPULib::matchStack(const uint8_t* rgbL, const uint8_t* rgbR, int rows, int cols, size_t step) {
    return stackImpl(rgbL, rgbR, rows, cols, step, nullptr, nullptr);
}
float*** MatchGPULib::matchStackPyramid(const uint8_t* rgbL, const uint8_t* rgbR, int rows, int cols, size_t step,
                                        float*** leftFoveated, float*** rightFoveated) {
    return stackImpl(rgbL, rgbR, rows, cols, step, leftFoveated, rightFoveated);
}

void ugsmFreeStack(float*** stack) {
    if (!stack) return;
    for (int i = 0; i < UGSM_MAX_LEVELS; i++) {
        if (!stack[i]) continue;
        for (int k = 0; k < 3; k++) free(stack[i][k]);
        free(stack[i]);
    }
    free(stack);
}

#ifdef UGSM_WITH_CV_BRIDGE
float** MatchGPULib::match(cv_bridge::CvImagePtr L, cv_bridge::CvImagePtr R, int fov) {
    return match((const uint8_t*)L->image.data, (const uint8_t*)R->image.data, (int)L->image.rows, (int)L->image.cols,
                 (size_t)L->image.step, fov);
}
int MatchGPULib::initStack(cv_bridge::CvImagePtr L, cv_bridge::CvImagePtr) {
    return initStack((int)L->image.rows, (int)L->image.cols);
}
float*** MatchGPULib::matchStack(cv_bridge::CvImagePtr L, cv_bridge::CvImagePtr R) {
    return matchStack((const uint8_t*)L->image.data, (const uint8_t*)R->image.data, (int)L->image.rows,
                      (int)L->image.cols, (size_t)L->image.step);
}
float*** MatchGPULib::matchStackPyramid(cv_bridge::CvImagePtr L, cv_bridge::CvImagePtr R, float*** lf, float*** rf) {
    return matchStackPyramid((const uint8_t*)L->image.data, (const uint8_t*)R->image.data, (int)L->image.rows,
                             (int)L->image.cols, (size_t)L->image.step, lf, rf);
}
#endif
