// ugsm_rowband.cu -- one pair split into row bands over the GPUs of a box, driven from C++ (SURVEY.md 8e;
// BASELINE.json configs[2]).  The reference has nothing here: MatchGPULib.cpp:303-403 is single-GPU.
//
// One process, one host thread per GPU.  Levels >= split_levels are computed replicated on every GPU (small, and
// it avoids tiny exchanges); the finest levels are split into contiguous row bands.  Per iteration a GPU runs K1
// on its band extended by S+1 rows (so K2 needs no exchange of K1's output) and K2 on its band, then the HALO rows
// next to each band boundary travel straight from one GPU's level planes into its neighbour's with
// cudaMemcpy2DAsync over NVLink peer access -- no staging buffer, no pack / unpack kernels, no Python.  Streams
// are ordered with events only: a copy waits for the iteration of both its ends, the next iteration waits for the
// copies into its halos; the host threads meet at a barrier once per iteration so that an event is always recorded
// before a neighbour waits on it.  Results are bit-identical to the single-GPU path.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/ugsm.h"

namespace {

constexpr int HALO = 16;  // rows exchanged per side and iteration: >= (S+1) + 3 for S = 10, with slack for the upsample

struct Barrier {  // reusable barrier of n host threads
    explicit Barrier(int n) : n_(n) {}
    void wait() {
        std::unique_lock<std::mutex> lk(m_);
        const int gen = gen_;
        if (++count_ == n_) {
            count_ = 0;
            gen_++;
            cv_.notify_all();
        } else {
            cv_.wait(lk, [&] { return gen != gen_; });
        }
    }
    std::mutex m_;
    std::condition_variable cv_;
    int n_, count_ = 0, gen_ = 0;
};

inline void band_rows(int height, int world, int rank, int* a, int* e) {
    *a = (int)(((long long)height * rank) / world);
    *e = (int)(((long long)height * (rank + 1)) / world);
}

// Fine rows [lo, hi) of the next level that rank must upsample itself (its band +- HALO); they may only read coarse
// rows that are valid on this rank (its coarse band +- HALO).  MGL:1526-1590: source row = floor((y + .5) * s).
bool upsample_rows(int ya_f, int yb_f, int h_f, int ya_c, int yb_c, int h_c, int* lo, int* hi) {
    const float sf = (float)(1 / 1.41421356);
    *lo = std::max(0, ya_f - HALO);
    *hi = std::min(h_f, yb_f + HALO);
    const int src_lo = (int)floorf(((float)*lo + 0.5f) * sf);
    const int src_hi = std::min(h_c - 1, (int)floorf(((float)(*hi - 1) + 0.5f) * sf));
    return src_lo >= std::max(0, ya_c - HALO) && src_hi <= std::min(h_c, yb_c + HALO) - 1;
}

struct Rank {
    ugsm_handle* h = nullptr;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;     // this rank's K1 + K2 of the current iteration have finished
    cudaEvent_t sent_dn = nullptr;  // its copy into the lower neighbour's top halo has finished
    cudaEvent_t sent_up = nullptr;  // its copy into the upper neighbour's bottom halo has finished
    int rc = UGSM_OK;
    char err[256] = "";
};

#define RCK(call)                                                                                      \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            snprintf(me.err, sizeof me.err, "%s failed: %s", #call, cudaGetErrorString(e_));          \
            me.rc = UGSM_ERR_CUDA;                                                                     \
        }                                                                                              \
    } while (0)
#define UCK(call)                                                           \
    do {                                                                    \
        int r_ = (call);                                                    \
        if (r_ != UGSM_OK && me.rc == UGSM_OK) {                            \
            snprintf(me.err, sizeof me.err, "%s", ugsm_last_error(me.h)); \
            me.rc = r_;                                                     \
        }                                                                   \
    } while (0)

thread_local char g_rb_error[512] = "";

}  // namespace

extern "C" {

const char* ugsm_rowband_last_error(void) { return g_rb_error; }

int ugsm_rowband_match(ugsm_handle* const* handles, const int* devices, int n, const uint8_t* const* d_rgbL,
                       const uint8_t* const* d_rgbR, int rows, int cols, size_t step, int split_levels, float* dH,
                       float* dV, float* dC, float* device_ms) {
    if (!handles || !devices || n < 1 || n > 64 || !d_rgbL || !d_rgbR || !dH || !dV || !dC || split_levels < 1) {
        snprintf(g_rb_error, sizeof g_rb_error, "ugsm_rowband_match: bad argument");
        return UGSM_ERR_INVALID;
    }
    std::vector<Rank> rk(n);
    for (int r = 0; r < n; r++) {
        rk[r].h = handles[r];
        rk[r].device = devices[r];
        if (!handles[r] || !d_rgbL[r] || !d_rgbR[r]) {
            snprintf(g_rb_error, sizeof g_rb_error, "ugsm_rowband_match: null handle or image for rank %d", r);
            return UGSM_ERR_INVALID;
        }
    }
    // peer access between neighbours (a no-op when both ranks are on the same device: single-GPU emulation)
    for (int r = 0; r < n; r++)
        for (int d : {r - 1, r + 1, 0}) {
            if (d < 0 || d >= n || devices[d] == devices[r]) continue;
            cudaSetDevice(devices[r]);
            cudaError_t e = cudaDeviceEnablePeerAccess(devices[d], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                snprintf(g_rb_error, sizeof g_rb_error, "no peer access from device %d to %d: %s", devices[r], devices[d],
                         cudaGetErrorString(e));
                cudaGetLastError();
                return UGSM_ERR_CUDA;
            }
            cudaGetLastError();
        }
    Barrier bar(n);
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    std::atomic<bool> failed{false};

    auto work = [&](int r) {
        Rank& me = rk[r];
        RCK(cudaSetDevice(me.device));
        RCK(cudaStreamCreateWithFlags(&me.stream, cudaStreamNonBlocking));
        RCK(cudaEventCreateWithFlags(&me.done, cudaEventDisableTiming));
        RCK(cudaEventCreateWithFlags(&me.sent_dn, cudaEventDisableTiming));
        RCK(cudaEventCreateWithFlags(&me.sent_up, cudaEventDisableTiming));
        if (r == 0) {
            RCK(cudaEventCreate(&t0));
            RCK(cudaEventCreate(&t1));
        }
        UCK(ugsm_set_stream(me.h, me.stream));
        if (r == 0) RCK(cudaEventRecord(t0, me.stream));
        // replicated part: pyramids, levels [levels-1 .. split_levels], upsample into level split_levels-1
        UCK(ugsm_band_begin(me.h, d_rgbL[r], d_rgbR[r], rows, cols, step, split_levels));
        ugsm_band_level prev{};
        int pya = 0, pyb = 0;
        if (me.rc != UGSM_OK) failed = true;
        for (int level = split_levels - 1; level >= 0; level--) {
            ugsm_band_level b{};
            UCK(ugsm_band_level_info(me.h, level, &b));
            int ya = 0, yb = 1;
            if (me.rc == UGSM_OK) band_rows(b.height, n, r, &ya, &yb);
            if (me.rc == UGSM_OK && n > 1 && yb - ya < HALO) {
                snprintf(me.err, sizeof me.err, "band of %d rows at level %d is smaller than the halo", yb - ya, level);
                me.rc = UGSM_ERR_INVALID;
            }
            if (level < split_levels - 1 && me.rc == UGSM_OK) {
                int lo, hi;
                if (!upsample_rows(ya, yb, b.height, pya, pyb, prev.height, &lo, &hi)) {
                    snprintf(me.err, sizeof me.err, "upsample into level %d would read rows outside the valid halo", level);
                    me.rc = UGSM_ERR_INVALID;
                } else {
                    // the coarse rows it reads include halos the neighbours' last copies wrote
                    if (r > 0) RCK(cudaStreamWaitEvent(me.stream, rk[r - 1].sent_dn, 0));
                    if (r < n - 1) RCK(cudaStreamWaitEvent(me.stream, rk[r + 1].sent_up, 0));
                    UCK(ugsm_band_upsample(me.h, level + 1, lo, hi));
                }
            }
            if (me.rc != UGSM_OK) failed = true;
            bar.wait();  // every rank knows this level's geometry (and whether anyone failed) before exchanging
            if (failed) break;
            const size_t row_bytes = (size_t)b.width * sizeof(float), pitch_bytes = (size_t)b.pitch * sizeof(float);
            for (int m = 1; m <= b.iters; m++) {
                // halos written by the neighbours' copies of the previous iteration (or level) must have landed
                if (m > 1) {
                    if (r > 0) RCK(cudaStreamWaitEvent(me.stream, rk[r - 1].sent_dn, 0));
                    if (r < n - 1) RCK(cudaStreamWaitEvent(me.stream, rk[r + 1].sent_up, 0));
                }
                UCK(ugsm_band_iteration(me.h, level, m, ya, yb, b.smooth + 1));
                RCK(cudaEventRecord(me.done, me.stream));
                bar.wait();  // all `done` events of this iteration are recorded
                // my boundary rows -> the neighbours' halos: after BOTH ends have finished the iteration (the
                // neighbour's K1 read the halo rows this copy overwrites)
                ugsm_band_level nb{};
                if (r < n - 1) {
                    UCK(ugsm_band_level_info(rk[r + 1].h, level, &nb));
                    RCK(cudaStreamWaitEvent(me.stream, rk[r + 1].done, 0));
                    for (int k = 0; k < 3; k++)
                        RCK(cudaMemcpy2DAsync(nb.state[k] + (size_t)(yb - HALO) * nb.pitch, pitch_bytes,
                                              b.state[k] + (size_t)(yb - HALO) * b.pitch, pitch_bytes, row_bytes, HALO,
                                              cudaMemcpyDeviceToDevice, me.stream));
                    RCK(cudaEventRecord(me.sent_dn, me.stream));
                }
                if (r > 0) {
                    UCK(ugsm_band_level_info(rk[r - 1].h, level, &nb));
                    RCK(cudaStreamWaitEvent(me.stream, rk[r - 1].done, 0));
                    for (int k = 0; k < 3; k++)
                        RCK(cudaMemcpy2DAsync(nb.state[k] + (size_t)ya * nb.pitch, pitch_bytes,
                                              b.state[k] + (size_t)ya * b.pitch, pitch_bytes, row_bytes, HALO,
                                              cudaMemcpyDeviceToDevice, me.stream));
                    RCK(cudaEventRecord(me.sent_up, me.stream));
                }
                bar.wait();  // all `sent` events are recorded before anyone waits on them
            }
            prev = b;
            pya = ya;
            pyb = yb;
        }
        // gather: every rank writes its band of the level-0 result straight into rank 0's dense output planes
        if (!failed && me.rc == UGSM_OK) {
            ugsm_band_level b{};
            UCK(ugsm_band_level_info(me.h, 0, &b));
            int ya, yb;
            band_rows(b.height, n, r, &ya, &yb);
            float* outs[3] = {dH, dV, dC};
            for (int k = 0; k < 3; k++)
                RCK(cudaMemcpy2DAsync(outs[k] + (size_t)ya * b.width, (size_t)b.width * sizeof(float),
                                      b.state[k] + (size_t)ya * b.pitch, (size_t)b.pitch * sizeof(float),
                                      (size_t)b.width * sizeof(float), yb - ya, cudaMemcpyDeviceToDevice, me.stream));
            RCK(cudaEventRecord(me.done, me.stream));
        }
        bar.wait();
        if (r == 0 && !failed) {
            for (int q = 1; q < n; q++) RCK(cudaStreamWaitEvent(me.stream, rk[q].done, 0));
            RCK(cudaEventRecord(t1, me.stream));
        }
        RCK(cudaStreamSynchronize(me.stream));
        bar.wait();  // nobody destroys an event a neighbour's stream may still wait on
        ugsm_set_stream(me.h, nullptr);
        if (r == 0 && device_ms && t0 && t1 && me.rc == UGSM_OK && !failed) RCK(cudaEventElapsedTime(device_ms, t0, t1));
        if (me.done) cudaEventDestroy(me.done);
        if (me.sent_dn) cudaEventDestroy(me.sent_dn);
        if (me.sent_up) cudaEventDestroy(me.sent_up);
        if (me.stream) cudaStreamDestroy(me.stream);
        if (r == 0) {
            if (t0) cudaEventDestroy(t0);
            if (t1) cudaEventDestroy(t1);
        }
    };

    int rc = UGSM_OK;
    // the ranks only start once every thread exists: a failed std::thread must not leave the others at a barrier
    std::atomic<int> gate{0};  // 0 = wait, 1 = go, -1 = abort
    std::vector<std::thread> pool;
    try {
        for (int r = 1; r < n; r++)
            pool.emplace_back([&, r] {
                while (gate.load() == 0) std::this_thread::yield();
                if (gate.load() > 0) work(r);
            });
    } catch (...) {
        gate = -1;
        for (auto& t : pool) t.join();
        snprintf(g_rb_error, sizeof g_rb_error, "ugsm_rowband_match: could not start the host threads");
        return UGSM_ERR_NOMEM;
    }
    gate = 1;
    work(0);
    for (auto& t : pool) t.join();
    for (int r = 0; r < n; r++)
        if (rk[r].rc != UGSM_OK && rc == UGSM_OK) {
            rc = rk[r].rc;
            snprintf(g_rb_error, sizeof g_rb_error, "rank %d: %s", r, rk[r].err);
        }
    return rc;
}

}  // extern "C"
