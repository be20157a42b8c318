// refcuda_main.cpp -- drives the UNMODIFIED reference MatchGPULib (built by oracle/refbuild/Makefile
// from /root/reference) on a raw RGB pair: refcuda <rows> <cols> <left.rgb> <right.rgb> <out.f32>
// Writes H, V, C (rows*cols float32 each).  Test infrastructure only.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "MatchGPULib.h"

int main(int argc, char** argv) {
    if (argc < 6) {
        fprintf(stderr, "usage: %s rows cols left.rgb right.rgb out.f32\n", argv[0]);
        return 2;
    }
    int rows = atoi(argv[1]), cols = atoi(argv[2]);
    size_t n = (size_t)rows * cols;
    std::vector<unsigned char> L(n * 3), R(n * 3);
    FILE* f = fopen(argv[3], "rb");
    if (!f || fread(L.data(), 1, n * 3, f) != n * 3) return 2;
    fclose(f);
    f = fopen(argv[4], "rb");
    if (!f || fread(R.data(), 1, n * 3, f) != n * 3) return 2;
    fclose(f);
    cv_bridge::CvImagePtr l(new cv_bridge::CvImage()), r(new cv_bridge::CvImage());
    l->image.rows = r->image.rows = rows;
    l->image.cols = r->image.cols = cols;
    l->image.step = r->image.step = (size_t)cols * 3;
    l->image.data = L.data();
    r->image.data = R.data();
    char* fake_argv[] = {argv[0], nullptr};
    MatchGPULib m(1, fake_argv);
    float** d = m.match(l, r, 0);
    f = fopen(argv[5], "wb");
    for (int k = 0; k < 3; k++) fwrite(d[k], sizeof(float), n, f);
    fclose(f);
    return 0;
}
