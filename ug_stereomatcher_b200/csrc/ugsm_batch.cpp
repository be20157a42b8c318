// ugsm_batch.cpp -- pair-level data-parallel streaming in one C++ process (SURVEY.md 8e way 1, BASELINE config 5:
// a batch of independent pairs over the GPUs of one box).  Worker threads each own a matcher handle (one CUDA
// stream) and pinned host buffers and pull pair indices from a shared counter; several workers per GPU keep the
// launch-bound coarse levels of one pair hidden behind the large kernels of the others.  No collective: pairs are
// independent.  Plain C-ABI only (include/ugsm.h), no CUDA toolchain needed to build this file.
//
//   ugsm_batch <rows> <cols> <pairs> [--devices 0,1,..] [--in-flight K] [--foveated] [--left l.rgb --right r.rgb]
//              [--list images.xml [--decode-threads T] [--ring-slots S]]
//
// --list: real images.  The TIFF pairs of an OpenCV image list (the reference's src/utils/input_images.xml format)
// are decoded by the library's feeder threads into a ring of page-locked buffers (ugsm_feeder_*, include/ugsm.h) and
// matched straight from there; rows / cols are taken from the first image (pass 0 0), <pairs> beyond the list repeat
// its last pair like the reference's reader.  Without --list and --left/--right a deterministic synthetic pair (hash
// noise, right = left shifted by 3 px) is used for every index.  Prints one JSON line: pairs/s over the whole batch (wall clock, copies included), per-pair latency,
// and a checksum of the last result (sum of H in double) that tests compare with the single-call path.
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ugsm.h"

static uint32_t hash32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7feb352dU;
    x ^= x >> 15;
    x *= 0x846ca68bU;
    x ^= x >> 16;
    return x;
}

struct Worker {
    int device = 0;
    ugsm_handle* h = nullptr;
    uint8_t *L = nullptr, *R = nullptr;
    float* out[3] = {nullptr, nullptr, nullptr};
    long done = 0;
    double busy_ms = 0, h_sum = 0;
    std::string err;
};

int main(int argc, char** argv) {
    if (argc < 4) {
        fprintf(stderr,
                "usage: %s rows cols pairs [--devices 0,1,..] [--in-flight K] [--foveated] [--left f --right f]\n",
                argv[0]);
        return 2;
    }
    int rows = atoi(argv[1]), cols = atoi(argv[2]);
    const long pairs = atol(argv[3]);
    std::vector<int> devices = {0};
    int in_flight = 4;
    bool fov = false;
    const char *lf = nullptr, *rf = nullptr, *list = nullptr;
    int decode_threads = 4, ring_slots = 0;
    for (int i = 4; i < argc; i++) {
        if (!strcmp(argv[i], "--devices") && i + 1 < argc) {
            devices.clear();
            for (char* t = strtok(argv[++i], ","); t; t = strtok(nullptr, ",")) devices.push_back(atoi(t));
        } else if (!strcmp(argv[i], "--in-flight") && i + 1 < argc) {
            in_flight = atoi(argv[++i]);
        } else if (!strcmp(argv[i], "--foveated")) {
            fov = true;
        } else if (!strcmp(argv[i], "--left") && i + 1 < argc) {
            lf = argv[++i];
        } else if (!strcmp(argv[i], "--right") && i + 1 < argc) {
            rf = argv[++i];
        } else if (!strcmp(argv[i], "--list") && i + 1 < argc) {
            list = argv[++i];
        } else if (!strcmp(argv[i], "--decode-threads") && i + 1 < argc) {
            decode_threads = atoi(argv[++i]);
        } else if (!strcmp(argv[i], "--ring-slots") && i + 1 < argc) {
            ring_slots = atoi(argv[++i]);
        }
    }
    ugsm_feeder* feeder = nullptr;
    if (list) {
        if (in_flight < 1 || devices.empty() || pairs < 1) {
            fprintf(stderr, "bad arguments\n");
            return 2;
        }
        if (ring_slots < 1) ring_slots = (int)devices.size() * in_flight + 2;
        if (ugsm_feeder_open(list, pairs, ring_slots, decode_threads, &feeder) != UGSM_OK) {
            fprintf(stderr, "cannot open image list %s: %s\n", list, ugsm_feeder_last_error(nullptr));
            return 2;
        }
        ugsm_feeder_info_t fi;
        ugsm_feeder_info(feeder, &fi);
        rows = fi.rows, cols = fi.cols;
    }
    if (rows < 1 || cols < 1 || pairs < 1 || in_flight < 1 || devices.empty()) {
        fprintf(stderr, "bad arguments\n");
        return 2;
    }
    ugsm_config cfg;
    ugsm_config_default(&cfg);
    ugsm_dims dims;
    if (ugsm_query_dims(&cfg, rows, cols, fov ? 1 : 0, &dims) != UGSM_OK) {
        fprintf(stderr, "invalid image size %dx%d\n", cols, rows);
        return 2;
    }
    const size_t n = (size_t)rows * cols, img_bytes = n * 3;
    const size_t out_floats = fov ? (size_t)dims.stack_rows * dims.fov_width : n;

    // the pair every worker matches
    std::vector<uint8_t> L0(img_bytes), R0(img_bytes);
    if (lf && rf) {
        FILE* f = fopen(lf, "rb");
        if (!f || fread(L0.data(), 1, img_bytes, f) != img_bytes) return 2;
        fclose(f);
        f = fopen(rf, "rb");
        if (!f || fread(R0.data(), 1, img_bytes, f) != img_bytes) return 2;
        fclose(f);
    } else {
        for (int y = 0; y < rows; y++)
            for (int x = 0; x < cols; x++)
                for (int c = 0; c < 3; c++) {
                    // smooth-ish texture: sum of two hash lattices, values kept inside [16, 240]
                    uint32_t a = hash32((uint32_t)((y >> 2) * 73856093) ^ (uint32_t)((x >> 2) * 19349663) ^ (c * 83492791u));
                    uint32_t b = hash32((uint32_t)(y * 2654435761u) ^ (uint32_t)(x * 40503u) ^ (c * 2246822519u));
                    L0[((size_t)y * cols + x) * 3 + c] = (uint8_t)(16 + (a % 160) + (b % 64));
                }
        for (int y = 0; y < rows; y++)
            for (int x = 0; x < cols; x++) {
                const int sx = x + 3 < cols ? x + 3 : cols - 1;  // a plain 3-pixel shift
                for (int c = 0; c < 3; c++) R0[((size_t)y * cols + x) * 3 + c] = L0[((size_t)y * cols + sx) * 3 + c];
            }
    }

    std::vector<Worker> w(devices.size() * (size_t)in_flight);
    for (size_t i = 0; i < w.size(); i++) {
        Worker& k = w[i];
        k.device = devices[i / in_flight];
        ugsm_config c = cfg;
        c.device = k.device;
        if (ugsm_create(&c, &k.h) != UGSM_OK) {
            fprintf(stderr, "ugsm_create on device %d failed: %s\n", k.device, ugsm_last_error(nullptr));
            return 3;
        }
        void* p[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
        const size_t sz[5] = {img_bytes, img_bytes, out_floats * 4, out_floats * 4, out_floats * 4};
        for (int b = 0; b < 5; b++)
            if (ugsm_host_alloc(sz[b], &p[b]) != UGSM_OK) {
                fprintf(stderr, "ugsm_host_alloc failed: %s\n", ugsm_last_error(nullptr));
                return 3;
            }
        k.L = (uint8_t*)p[0], k.R = (uint8_t*)p[1];
        for (int b = 0; b < 3; b++) k.out[b] = (float*)p[2 + b];
        memcpy(k.L, L0.data(), img_bytes);
        memcpy(k.R, R0.data(), img_bytes);
    }
    auto run_pair = [&](Worker& k, const uint8_t* L, const uint8_t* R) -> int {
        return fov ? ugsm_match_foveated(k.h, L, R, rows, cols, (size_t)cols * 3, k.out[0], k.out[1], k.out[2], nullptr,
                                         nullptr)
                   : ugsm_match_full(k.h, L, R, rows, cols, (size_t)cols * 3, k.out[0], k.out[1], k.out[2]);
    };
    auto run_one = [&](Worker& k) -> int { return run_pair(k, k.L, k.R); };
    // warm-up: buffers, graphs and clocks, one pair per worker, concurrently
    {
        std::vector<std::thread> t;
        for (auto& k : w)
            t.emplace_back([&]() {
                if (run_one(k) != UGSM_OK) k.err = ugsm_last_error(k.h);
            });
        for (auto& th : t) th.join();
        for (auto& k : w)
            if (!k.err.empty()) {
                fprintf(stderr, "warm-up failed on device %d: %s\n", k.device, k.err.c_str());
                return 3;
            }
    }
    std::atomic<long> next(0);
    const auto t0 = std::chrono::steady_clock::now();
    {
        std::vector<std::thread> t;
        for (auto& k : w)
            t.emplace_back([&]() {
                for (;;) {
                    ugsm_feeder_pair fp;
                    if (feeder) {  // real images: the next decoded pair of the list, matched from its ring slot
                        const int rc = ugsm_feeder_next(feeder, &fp);
                        if (rc == UGSM_FEEDER_END) break;
                        if (rc != UGSM_OK) {
                            k.err = ugsm_feeder_last_error(feeder);
                            break;
                        }
                    } else {
                        const long i = next.fetch_add(1);
                        if (i >= pairs) break;
                    }
                    const auto a = std::chrono::steady_clock::now();
                    const int mrc = feeder ? run_pair(k, fp.left, fp.right) : run_one(k);
                    if (feeder) ugsm_feeder_release(feeder, fp.slot);
                    if (mrc != UGSM_OK) {
                        k.err = ugsm_last_error(k.h);
                        break;
                    }
                    k.busy_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count();
                    k.done++;
                }
            });
        for (auto& th : t) th.join();
    }
    const double wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    long done = 0;
    double busy = 0;
    for (auto& k : w) {
        if (!k.err.empty()) {
            fprintf(stderr, "match failed on device %d: %s\n", k.device, k.err.c_str());
            return 3;
        }
        done += k.done;
        busy += k.busy_ms;
    }
    double h_sum = 0;
    for (size_t i = 0; i < out_floats; i++) h_sum += (double)w[0].out[0][i];
    printf("{\"tool\": \"ugsm_batch\", \"mode\": \"%s\", \"rows\": %d, \"cols\": %d, \"pairs\": %ld, \"gpus\": %zu, "
           "\"in_flight_per_gpu\": %d, \"wall_ms\": %.3f, \"pairs_per_s\": %.3f, \"ms_per_pair_latency\": %.3f, "
           "\"h_sum\": %.9e",
           fov ? "foveated" : "full", rows, cols, done, devices.size(), in_flight, wall_ms, done / (wall_ms * 1e-3),
           done ? busy / done : 0.0, h_sum);
    if (feeder) {
        ugsm_feeder_info_t fi;
        ugsm_feeder_info(feeder, &fi);
        printf(", \"input\": \"tiff list\", \"list_images\": %lld, \"decode_threads\": %d, \"ring_slots\": %d, "
               "\"ring_pinned\": %d, \"decode_ms_per_pair\": %.3f",
               (long long)fi.list_images, decode_threads, fi.slots, fi.pinned, done ? fi.decode_ms / done : 0.0);
        ugsm_feeder_close(feeder);
    }
    printf("}\n");
    for (auto& k : w) {
        ugsm_destroy(k.h);
        ugsm_host_free(k.L);
        ugsm_host_free(k.R);
        for (int b = 0; b < 3; b++) ugsm_host_free(k.out[b]);
    }
    return 0;
}
