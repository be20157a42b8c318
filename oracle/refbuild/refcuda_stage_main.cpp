// refcuda_stage_main.cpp -- calls individual public methods of the UNMODIFIED reference MatchGPULib on
// raw float planes, so that each stage can be compared with the oracle at sizes where the reference's
// partial-tile accidents cannot trigger (W % 128 == 0, H % 64 == 0).  Test infrastructure only.
//   refcuda_stage level  W H level in.f32(L3,R3,H,V,C) out.f32(H,V,C)     -> MatchGPULib::matchlevel
//   refcuda_stage conv   W H in.f32(3 planes) out.f32                     -> MatchGPULib::convolutionGPU
//   refcuda_stage decim  W H scale W2 H2 in.f32(3 planes) out.f32         -> MatchGPULib::subsampledimageGPU
//   refcuda_stage updisp W H W2 H2 in.f32(3 planes) out.f32               -> MatchGPULib::subsampleDisp
//   refcuda_stage hier   W H fl in.f32(stack: 3 x fl x fovH x fovW) out.f32 -> MatchGPULib::hierarchicalDisparity
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "MatchGPULib.h"
#include "MatchLib_common.h"

static std::vector<float> rd(const char* path, size_t n) {
    std::vector<float> v(n);
    FILE* f = fopen(path, "rb");
    if (!f || fread(v.data(), 4, n, f) != n) {
        fprintf(stderr, "cannot read %zu floats from %s\n", n, path);
        exit(2);
    }
    fclose(f);
    return v;
}
static void wr(const char* path, float** planes, int np, size_t n) {
    FILE* f = fopen(path, "wb");
    for (int k = 0; k < np; k++) fwrite(planes[k], 4, n, f);
    fclose(f);
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    std::string mode = argv[1];
    char a1[] = "x";
    char* fake_argv[] = {argv[0], a1, (mode == "hier" && argc > 4) ? argv[4] : nullptr, nullptr};
    MatchGPULib m(fake_argv[2] ? 3 : 1, fake_argv);  // argv[2] = foveatelevel (MatchGPULib.cpp:259-264)
    float gk[KERNEL_LENGTH], ak[KERNEL_LENGTH] = {0.0, 0.3333, 0.3333, 0.3333, 0.0};
    m.gaussiankernel(gk);
    setConvolutionKernel(gk);
    setConvolutionAverageKernel(ak);
    if (mode == "level") {
        int W = atoi(argv[2]), H = atoi(argv[3]), level = atoi(argv[4]);
        size_t n = (size_t)W * H;
        std::vector<float> in = rd(argv[5], 9 * n);
        float *L[3], *R[3], *D[3];
        for (int k = 0; k < 3; k++) {
            L[k] = in.data() + k * n;
            R[k] = in.data() + (3 + k) * n;
            D[k] = in.data() + (6 + k) * n;
        }
        float** out = m.matchlevel(L, R, D, 3, W, H, level, 1, 1.0f);
        wr(argv[6], out, 3, n);
    } else if (mode == "conv") {
        int W = atoi(argv[2]), H = atoi(argv[3]);
        size_t n = (size_t)W * H;
        std::vector<float> in = rd(argv[4], 3 * n);
        float* P[3] = {in.data(), in.data() + n, in.data() + 2 * n};
        float** out = m.convolutionGPU(P, 3, W, H);
        wr(argv[5], out, 3, n);
    } else if (mode == "decim") {
        int W = atoi(argv[2]), H = atoi(argv[3]);
        float s = (float)atof(argv[4]);
        int W2 = atoi(argv[5]), H2 = atoi(argv[6]);
        size_t n = (size_t)W * H;
        std::vector<float> in = rd(argv[7], 3 * n);
        float* P[3] = {in.data(), in.data() + n, in.data() + 2 * n};
        float** out = m.subsampledimageGPU(P, s, 3, W, H, W2, H2);
        wr(argv[8], out, 3, (size_t)W2 * H2);
    } else if (mode == "updisp") {
        int W = atoi(argv[2]), H = atoi(argv[3]), W2 = atoi(argv[4]), H2 = atoi(argv[5]);
        size_t n = (size_t)W * H;
        std::vector<float> in = rd(argv[6], 3 * n);
        float* P[3] = {in.data(), in.data() + n, in.data() + 2 * n};
        float** out = m.subsampleDisp(P, (float)(1 / SCALE), 3, W, H, W2, H2);
        wr(argv[7], out, 3, (size_t)W2 * H2);
    } else if (mode == "hier") {
        int W = atoi(argv[2]), H = atoi(argv[3]), fl = atoi(argv[4]);
        int w[16], h[16];
        w[0] = W;
        h[0] = H;
        for (int i = 0; i < fl; i++) {
            w[i + 1] = w[i] / SCALE;
            h[i + 1] = h[i] / SCALE;
        }
        int fw = w[fl - 1], fh = h[fl - 1];
        size_t n = (size_t)fw * fh;
        std::vector<float> in = rd(argv[5], 3 * fl * n);
        std::vector<float**> lv(fl);
        std::vector<std::vector<float*>> pl(fl, std::vector<float*>(3));
        for (int i = 0; i < fl; i++) {
            for (int k = 0; k < 3; k++) pl[i][k] = in.data() + ((size_t)k * fl + i) * n;
            lv[i] = pl[i].data();
        }
        float** out = m.hierarchicalDisparity(nullptr, lv.data(), 3, W, H);
        wr(argv[6], out, 3, (size_t)W * H);
    } else {
        return 2;
    }
    return 0;
}
