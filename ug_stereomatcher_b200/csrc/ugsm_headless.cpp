// ugsm_headless.cpp -- the hot path without ROS: reads a raw RGB8 pair, runs the MatchGPULib facade
// exactly as the ROS node does (UG_GPU_matcher.cpp:160-185, 423), writes the float planes.
//   ugsm_headless <rows> <cols> <left.rgb> <right.rgb> <out.f32> [--foveated] [--fovea-level N] [--repeat K]
// full mode:     out = H, V, C planes (rows*cols floats each)
// foveated mode: out = stack H, V, C ((foveatelevel*fovH) x fovW floats each, level-major rows,
//                UG_GPU_matcher.cpp:293-303)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/MatchGPULib.h"

int main(int argc, char** argv) {
    if (argc < 6) {
        fprintf(stderr, "usage: %s rows cols left.rgb right.rgb out.f32 [--foveated] [--fovea-level N] [--repeat K]\n",
                argv[0]);
        return 2;
    }
    int rows = atoi(argv[1]), cols = atoi(argv[2]);
    bool fov = false;
    int repeat = 1;
    std::string flevel = "7";
    for (int i = 6; i < argc; i++) {
        if (!strcmp(argv[i], "--foveated")) fov = true;
        if (!strcmp(argv[i], "--fovea-level") && i + 1 < argc) flevel = argv[++i];
        if (!strcmp(argv[i], "--repeat") && i + 1 < argc) repeat = atoi(argv[++i]);
    }
    size_t n = (size_t)rows * cols;
    std::vector<unsigned char> L(n * 3), R(n * 3);
    FILE* f = fopen(argv[3], "rb");
    if (!f || fread(L.data(), 1, n * 3, f) != n * 3) return 2;
    fclose(f);
    f = fopen(argv[4], "rb");
    if (!f || fread(R.data(), 1, n * 3, f) != n * 3) return 2;
    fclose(f);

    char a0[] = "ugsm_headless", a1[] = "x";
    std::vector<char> a2(flevel.begin(), flevel.end());
    a2.push_back(0);
    char* margv[] = {a0, a1, a2.data(), nullptr};
    MatchGPULib mgpu(3, margv);  // argv[2] = fovea level, as in the node
    mgpu.setFoveated(fov ? 1 : 0);
    FILE* out = nullptr;
    for (int r = 0; r < repeat; r++) {
        auto t0 = std::chrono::steady_clock::now();
        if (!fov) {
            float** d = mgpu.match(L.data(), R.data(), rows, cols, (size_t)cols * 3, 0);
            if (!d) {
                fprintf(stderr, "match failed: %s\n", mgpu.lastError());
                return 3;
            }
            if (r == repeat - 1) {
                out = fopen(argv[5], "wb");
                for (int k = 0; k < 3; k++) fwrite(d[k], sizeof(float), n, out);
                fclose(out);
            }
            for (int k = 0; k < 3; k++) free(d[k]);  // UG_GPU_matcher.cpp:487-489
            free(d);
        } else {
            mgpu.initStack(rows, cols);
            float*** s = mgpu.matchStack(L.data(), R.data(), rows, cols, (size_t)cols * 3);
            if (!s) {
                fprintf(stderr, "matchStack failed: %s\n", mgpu.lastError());
                return 3;
            }
            if (r == repeat - 1) {
                size_t fn = (size_t)mgpu.getFoveaWidth() * mgpu.getFoveaHeight();
                out = fopen(argv[5], "wb");
                for (int k = 0; k < 3; k++)
                    for (int lv = 0; lv < mgpu.getFoveateLevel(); lv++) fwrite(s[lv][k], sizeof(float), fn, out);
                fclose(out);
            }
            ugsmFreeStack(s);
        }
        double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        fprintf(stderr, "pair %d: %.2f ms wall, %.2f ms on device%s\n", r, ms, mgpu.lastDeviceMs(),
                fov ? " (foveated)" : "");
    }
    return 0;
}
