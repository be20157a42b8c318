// Minimal stand-in for the CUDA samples' helper_functions.h: the StopWatch timers only.
#pragma once
#include <chrono>
struct StopWatchInterface {
    std::chrono::steady_clock::time_point t0;
    double acc_ms = 0;
    bool running = false;
};
static inline bool sdkCreateTimer(StopWatchInterface** t) { *t = new StopWatchInterface(); return true; }
static inline bool sdkDeleteTimer(StopWatchInterface** t) { delete *t; *t = nullptr; return true; }
static inline bool sdkResetTimer(StopWatchInterface** t) { (*t)->acc_ms = 0; (*t)->running = false; return true; }
static inline bool sdkStartTimer(StopWatchInterface** t) { (*t)->t0 = std::chrono::steady_clock::now(); (*t)->running = true; return true; }
static inline bool sdkStopTimer(StopWatchInterface** t) {
    if ((*t)->running) (*t)->acc_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - (*t)->t0).count();
    (*t)->running = false;
    return true;
}
static inline float sdkGetTimerValue(StopWatchInterface** t) { return (float)(*t)->acc_ms; }
