/*
 * ugsm_oracle.c -- CPU oracle (test infrastructure only; see ugsm_oracle.h).
 *
 * Reference files cited below, relative to the reference checkout:
 *   ML  = src/gpu_matcher/MatchLib.cu
 *   MGL = src/gpu_matcher/MatchGPULib.cpp
 *   GOLD= src/gpu_matcher/convolutionSeparable_gold.cpp
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -mfma (oracle/Makefile).  -ffp-contract=off
 * keeps gcc from fusing anything on its own; the fmaf()/fma() calls below are the
 * multiply-adds nvcc fuses in the reference kernels (FFMA/DFMA in their SASS).
 *
 * Deliberate deviations from reference *accidents* (not replicated):
 *   - coarsest-level H,V,C start at 0 instead of uninitialised malloc (MGL:1247);
 *   - zero-border convolution is clean (no partial-tile wrap reads / overflow writes,
 *     ML:97-100,121-137,220-223,244-260) -- GOLD:20-75 defines the in-image result;
 *   - texture fetches are explicit floorf + clamp (point filter, clamp addressing,
 *     unnormalised coordinates: the defaults of the texture references at ML:56-60).
 */
#include "ugsm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define SCALE_D 1.41421356 /* MatchLib_common.h:15 */

void ugsmo_config_default(ugsmo_config* cfg) {
    cfg->levels = 14;            /* MAX_LEVEL, MatchLib_common.h:13 */
    cfg->fovea_level = 7;        /* MGL:263 */
    cfg->iters_cutoff = 22;      /* levelcutoff, MGL:37 */
    cfg->smooth_passes = 5;      /* smoothtime, MGL:38 */
    cfg->smooth_passes_fine = 10; /* MGL:2259-2261 */
    cfg->max_iters = 0;
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arms of bench.py ask for the host cores explicitly */
void ugsmo_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int ugsmo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Texture fetch index: point filter, clamp addressing, unnormalised coordinate. */
static inline int tex_index(float c, int n) {
    float f = floorf(c);
    if (!(f >= 0.0f)) return 0; /* also catches NaN */
    if (f > (float)(n - 1)) return n - 1;
    return (int)f;
}
static inline int clampi(int v, int n) { return v < 0 ? 0 : (v > n - 1 ? n - 1 : v); }

/* MGL:761-775: the analytic taps (MGL:742-759) are overwritten by five literals, then
 * renormalised by their float sum in index order.  Box taps: MGL:344-348. */
void ugsmo_taps(float gauss[5], float box[5]) {
    gauss[0] = 0.0816475;
    gauss[1] = 0.218507;
    gauss[2] = 0.303281;
    gauss[3] = 0.218507;
    gauss[4] = 0.0816475;
    float kern = 0;
    for (int i = 0; i < 5; i++) kern = kern + gauss[i];
    for (int i = 0; i < 5; i++) gauss[i] = gauss[i] / kern;
    box[0] = 0.0;
    box[1] = 0.3333;
    box[2] = 0.3333;
    box[3] = 0.3333;
    box[4] = 0.0;
}

/* MGL:1224-1228 (same recurrence at MGL:1084-1100): int <- int / double, truncation. */
void ugsmo_level_dims(int cols, int rows, int levels, int* w, int* h) {
    w[0] = cols;
    h[0] = rows;
    for (int i = 0; i + 1 < levels; i++) {
        w[i + 1] = w[i] / SCALE_D;
        h[i + 1] = h[i] / SCALE_D;
    }
}

/* MGL:1741 with level = MAX_LEVEL-1-i (MGL:1266): mi = (i>5) ? levelcutoff : (i+1)*2. */
int ugsmo_level_iters(const ugsmo_config* cfg, int i) {
    int mi = (i > 5) ? cfg->iters_cutoff : ((i + 1) * 2);
    if (cfg->max_iters > 0 && mi > cfg->max_iters) mi = cfg->max_iters;
    return mi;
}
/* MGL:2257-2261: level > 11  <=>  i < 2. */
int ugsmo_level_smooth(const ugsmo_config* cfg, int i) {
    return (i < 2) ? cfg->smooth_passes_fine : cfg->smooth_passes;
}

/* MGL:2299-2306.  Evaluated after iteration m; the value applies from iteration m+1. */
float ugsmo_threshold_after(int mi, int m, float T) {
    float threshold = T;
    if (m % 2 == 0) {
        if ((mi / 2 - m / 2) < 7) {
            threshold = ((mi / 2 - m / 2) - 1) * ((1 - 0.1) / (mi / 2 - 1.0)) + 0.1;
        } else {
            threshold = 1.0;
        }
    }
    return threshold;
}

/* MGL:332-338 */
void ugsmo_unpack_rgb8(const uint8_t* rgb, int rows, int cols, size_t step, float* planes3) {
    size_t n = (size_t)rows * cols;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < rows; i++)
        for (int j = 0; j < cols; j++)
            for (int k = 0; k < 3; k++) planes3[k * n + (size_t)i * cols + j] = (float)rgb[i * step + (size_t)j * 3 + k];
}

/* ML:131-136 / GOLD:31-40: sum = 0; for j=-2..2: sum += taps[2-j]*src[x+j] (FFMA chain in
 * the reference SASS), zero outside the image. */
void ugsmo_conv_rows_zero(const float* src, float* dst, int w, int h, const float taps[5]) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++) {
        const float* s = src + (size_t)y * w;
        float* d = dst + (size_t)y * w;
        for (int x = 0; x < w; x++) {
            float sum = 0;
            for (int j = -2; j <= 2; j++) {
                int xx = x + j;
                float v = (xx >= 0 && xx < w) ? s[xx] : 0.0f;
                sum = fmaf(taps[2 - j], v, sum);
            }
            d[x] = sum;
        }
    }
}

/* ML:254-259 / GOLD:61-70 */
void ugsmo_conv_cols_zero(const float* src, float* dst, int w, int h, const float taps[5]) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++) {
        float* d = dst + (size_t)y * w;
        for (int x = 0; x < w; x++) {
            float sum = 0;
            for (int j = -2; j <= 2; j++) {
                int yy = y + j;
                float v = (yy >= 0 && yy < h) ? src[(size_t)yy * w + x] : 0.0f;
                sum = fmaf(taps[2 - j], v, sum);
            }
            d[x] = sum;
        }
    }
}

/* ML:1484-1491 (c_Kernel) and ML:1616-1623 (average): tex2D(x+k, y) * taps[2-k], clamp addressing. */
void ugsmo_conv_rows_clamp(const float* src, float* dst, int w, int h, const float taps[5]) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++) {
        const float* s = src + (size_t)y * w;
        float* d = dst + (size_t)y * w;
        for (int x = 0; x < w; x++) {
            float sum = 0;
            for (int k = -2; k <= 2; k++) sum = fmaf(s[clampi(x + k, w)], taps[2 - k], sum);
            d[x] = sum;
        }
    }
}

/* ML:1551-1558 and ML:1683-1690 */
void ugsmo_conv_cols_clamp(const float* src, float* dst, int w, int h, const float taps[5]) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++) {
        float* d = dst + (size_t)y * w;
        for (int x = 0; x < w; x++) {
            float sum = 0;
            for (int k = -2; k <= 2; k++) sum = fmaf(src[(size_t)clampi(y + k, h) * w + x], taps[2 - k], sum);
            d[x] = sum;
        }
    }
}

/* ML:311-339: dst[iy][ix] = tex2D(src, (ix+.5)*s, (iy+.5)*s); the guard at ML:330 always
 * holds for the dims MGL:1084-1095 passes (w2 <= w/s), so every pixel is written. */
void ugsmo_decimate(const float* src, int w, int h, float scale, float* dst, int w2, int h2) {
#pragma omp parallel for schedule(static)
    for (int iy = 0; iy < h2; iy++) {
        float y = (float)iy + 0.5f;
        int sy = tex_index(y * scale, h);
        for (int ix = 0; ix < w2; ix++) {
            float x = (float)ix + 0.5f;
            int sx = tex_index(x * scale, w);
            dst[(size_t)iy * w2 + ix] = src[(size_t)sy * w + sx];
        }
    }
}

/* ML:372-401 with scalefactor = (float)(1/SCALE) (MGL:1222): nearest upsample, value * SCALE in double. */
void ugsmo_upsample_disp(const float* src, int w, int h, float* dst, int w2, int h2) {
    const float scalefactor = 1 / SCALE_D;
#pragma omp parallel for schedule(static)
    for (int iy = 0; iy < h2; iy++) {
        float y = (float)iy + 0.5f;
        int sy = tex_index(y * scalefactor, h);
        for (int ix = 0; ix < w2; ix++) {
            float x = (float)ix + 0.5f;
            int sx = tex_index(x * scalefactor, w);
            float v = src[(size_t)sy * w + sx];
            dst[(size_t)iy * w2 + ix] = SCALE_D * v;
        }
    }
}

/* ML:499-520: dst = tex2D(R, x + dH, y + dV), float adds, point filter, clamp. */
void ugsmo_warp(const float* R, const float* H, const float* V, int w, int h, float* dst, int32_t* srcx,
                int32_t* srcy) {
#pragma omp parallel for schedule(static)
    for (int iy = 0; iy < h; iy++) {
        float y = (float)iy + 0.5f;
        for (int ix = 0; ix < w; ix++) {
            float x = (float)ix + 0.5f;
            size_t p = (size_t)iy * w + ix;
            int sx = tex_index(x + H[p], w);
            int sy = tex_index(y + V[p], h);
            if (dst) dst[p] = R[(size_t)sy * w + sx];
            if (srcx) srcx[p] = sx;
            if (srcy) srcy[p] = sy;
        }
    }
}

/* ML:790-843.  Mixed float/double exactly as the kernel (and its SASS) computes it:
 *   b1 = (r-l)/2, c1 = r-(c+b1)                          float
 *   h  = (float)((-b1*0.5)/c1)                           double divide
 *   h  = (float)min(T, max(h, 0.0-T))                    double compare, exact
 *   c* = fma(h, fma(c1,h,b1), c)                         two FFMA
 *   if c* > 1: d = c*-c; if ((double)d > 1e-10) h = (float)(h*((1.0-c)/d)); conf = 1
 *   else conf = (float)fma(c*, 0.3, 0.7)                 one DFMA
 *   c1 >= 0 or NaN: h = 0, conf = 0.4 */
void ugsmo_poly(float l, float c, float r, float T, float* offset, float* conf, uint8_t* branch) {
    float b1 = (r - l) / 2;
    float c1 = r - (c + b1);
    if (c1 < 0) {
        float hor = (float)(((double)(-b1) * 0.5) / (double)c1);
        hor = (float)fmin((double)T, fmax((double)hor, 0.0 - (double)T));
        float cstar = fmaf(hor, fmaf(c1, hor, b1), c);
        if (cstar > 1.0f) {
            float d = cstar - c;
            if ((double)d > 1e-10) hor = (float)((double)hor * ((1.0 - (double)c) / (double)d));
            *offset = hor;
            *conf = 1.0f;
            if (branch) *branch = 2;
        } else {
            *offset = hor;
            *conf = (float)fma((double)cstar, 0.3, 0.7);
            if (branch) *branch = 1;
        }
    } else {
        *offset = 0.0f;
        *conf = (float)0.4;
        if (branch) *branch = 0;
    }
}

void ugsmo_poly_n(const float* l, const float* c, const float* r, int n, float T, float* offset, float* conf,
                  uint8_t* branch) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) ugsmo_poly(l[i], c[i], r[i], T, &offset[i], &conf[i], &branch[i]);
}

/* ML:1092-1145: five-point confidence-weighted average, centre, x-1, x+1, y-1, y+1 in
 * that order; row 0 and column 0 are not written (caller pre-fills dst with src). */
void ugsmo_smooth_pass(const float* src, const float* conf, float* dst, int w, int h) {
#pragma omp parallel for schedule(static)
    for (int iy = 1; iy < h; iy++) {
        int yp = clampi(iy + 1, h);
        for (int ix = 1; ix < w; ix++) {
            int xp = clampi(ix + 1, w);
            size_t pc = (size_t)iy * w + ix, pl = pc - 1, pr = (size_t)iy * w + xp, pu = pc - w,
                   pd = (size_t)yp * w + ix;
            float sumDisp = 0, sumCorr = 0;
            sumDisp = fmaf(src[pc], conf[pc], sumDisp);
            sumCorr = sumCorr + conf[pc];
            sumDisp = fmaf(src[pl], conf[pl], sumDisp);
            sumCorr = sumCorr + conf[pl];
            sumDisp = fmaf(src[pr], conf[pr], sumDisp);
            sumCorr = sumCorr + conf[pr];
            sumDisp = fmaf(src[pu], conf[pu], sumDisp);
            sumCorr = sumCorr + conf[pu];
            sumDisp = fmaf(src[pd], conf[pd], sumDisp);
            sumCorr = sumCorr + conf[pd];
            dst[pc] = sumDisp / sumCorr;
        }
    }
}

/* MGL:1033-1125.  B_i = cols(rows(level i)) zero border (MGL:866-960); level 1 =
 * decimate(B_0, (float)SCALE) (MGL:1082-1087); level i+2 = decimate(B_i, 2.0f) for
 * i < levels-2 (MGL:1088-1096, scalefactors = (int)(SCALE*SCALE+0.5)). */
void ugsmo_pyramid_plane(const float* plane0, int cols, int rows, int levels, float* const* out) {
    int w[UGSMO_MAX_LEVELS], h[UGSMO_MAX_LEVELS];
    float g[5], a[5];
    ugsmo_taps(g, a);
    ugsmo_level_dims(cols, rows, levels, w, h);
    size_t n0 = (size_t)cols * rows;
    float* tmp = (float*)malloc(n0 * sizeof(float));
    float* blur = (float*)malloc(n0 * sizeof(float));
    memcpy(out[0], plane0, n0 * sizeof(float));
    for (int i = 0; i < levels; i++) {
        int feeds1 = (i == 0 && levels > 1);
        int feeds2 = (i + 2 < levels);
        if (!feeds1 && !feeds2) continue; /* reference blurs levels 12,13 and discards them */
        ugsmo_conv_rows_zero(out[i], tmp, w[i], h[i], g);
        ugsmo_conv_cols_zero(tmp, blur, w[i], h[i], g);
        if (feeds1) {
            float s = SCALE_D;
            ugsmo_decimate(blur, w[i], h[i], s, out[1], w[1], h[1]);
        }
        if (feeds2) {
            float s = 0.000 + (int)(SCALE_D * SCALE_D + 0.5);
            ugsmo_decimate(blur, w[i], h[i], s, out[i + 2], w[i + 2], h[i + 2]);
        }
    }
    free(tmp);
    free(blur);
}

/* One iteration of MGL:1743-2250 ("usingMoreGPUMemory == 1" branch, MGL:35). */
void ugsmo_correlate(const float* L3, const float* R3, const float* H, const float* V, const float* C, int w, int h,
                     float T, int first_coarsest, float* Hn, float* Vn, float* Cn, int32_t* srcx, int32_t* srcy,
                     uint8_t* branch_h, uint8_t* branch_v) {
    size_t n = (size_t)w * h;
    float g[5], a[5];
    ugsmo_taps(g, a);
    float* Rw = (float*)malloc(n * sizeof(float));
    float* U = (float*)malloc(n * sizeof(float));
    float* D = (float*)malloc(n * sizeof(float));
    float* P = (float*)malloc(n * sizeof(float));
    float* t1 = (float*)malloc(n * sizeof(float));
    float* N = (float*)malloc(n * sizeof(float));
    float* A[5];
    for (int t = 0; t < 5; t++) A[t] = (float*)malloc(n * sizeof(float));
    /* MGL:1677: {(-1,0),(+1,0),(0,-1),(0,+1),(0,0)} as (x,y) */
    static const int mvx[5] = {-1, 1, 0, 0, 0};
    static const int mvy[5] = {0, 0, -1, 1, 0};

    for (int ch = 0; ch < 3; ch++) {
        const float* L = L3 + ch * n;
        const float* R = R3 + ch * n;
        /* MGL:1792 warp by the iteration-start H,V (disp0/disp1, MGL:1771-1772) */
        ugsmo_warp(R, H, V, w, h, Rw, ch == 0 ? srcx : NULL, ch == 0 ? srcy : NULL);
        /* MGL:1809-1901: U = colT(rowT(L^2)), D = colT(rowT(Rw^2)), Gaussian taps, clamp */
#pragma omp parallel for schedule(static)
        for (size_t p = 0; p < n; p++) P[p] = L[p] * L[p]; /* ML:570 */
        ugsmo_conv_rows_clamp(P, t1, w, h, g);
        ugsmo_conv_cols_clamp(t1, U, w, h, g);
#pragma omp parallel for schedule(static)
        for (size_t p = 0; p < n; p++) P[p] = Rw[p] * Rw[p];
        ugsmo_conv_rows_clamp(P, t1, w, h, g);
        ugsmo_conv_cols_clamp(t1, D, w, h, g);

        for (int t = 0; t < 5; t++) {
            /* ML:622-624 CompareMove: P = L(x,y) * Rw(clamp(x+tx), clamp(y+ty)) */
#pragma omp parallel for schedule(static)
            for (int y = 0; y < h; y++) {
                int yy = clampi(y + mvy[t], h);
                for (int x = 0; x < w; x++) {
                    int xx = clampi(x + mvx[t], w);
                    P[(size_t)y * w + x] = L[(size_t)y * w + x] * Rw[(size_t)yy * w + xx];
                }
            }
            /* MGL:1932-1945: N = colS(rowS(P)) zero border */
            ugsmo_conv_rows_zero(P, t1, w, h, g);
            ugsmo_conv_cols_zero(t1, N, w, h, g);
            /* ML:681-687 MoveCorrelation + MGL:2033-2070 RGB accumulation */
#pragma omp parallel for schedule(static)
            for (int y = 0; y < h; y++) {
                int yy = clampi(y + mvy[t], h);
                for (int x = 0; x < w; x++) {
                    int xx = clampi(x + mvx[t], w);
                    size_t p = (size_t)y * w + x;
                    float src = N[p], warpl = U[p], warpr = D[(size_t)yy * w + xx];
                    float rho = (src * src) / (warpl * warpr);
                    if (rho > 1) rho = 1.0;
                    if (rho < 0) rho = 0.0;
                    if (ch == 0)
                        A[t][p] = rho;
                    else if (ch == 1)
                        A[t][p] = rho + A[t][p]; /* ML:953 */
                    else
                        A[t][p] = (A[t][p] + rho) / 3.0f; /* ML:1393 */
                }
            }
        }
    }

#pragma omp parallel for schedule(static)
    for (size_t p = 0; p < n; p++) {
        float hH, cH, hV, cV;
        /* MGL:2129: l = left move, r = right move; MGL:2143: l = up move, r = down move */
        ugsmo_poly(A[0][p], A[4][p], A[1][p], T, &hH, &cH, branch_h ? &branch_h[p] : NULL);
        ugsmo_poly(A[2][p], A[4][p], A[3][p], T, &hV, &cV, branch_v ? &branch_v[p] : NULL);
        float cn = cV * cH;  /* ML:900 */
        float hn = hH + H[p]; /* ML:953; the x1 "scale" (ML:1057, int m <- 1.0f) is identity */
        float vn = hV + V[p];
        float c;
        if (first_coarsest) { /* MGL:2223 */
            c = cn;
        } else { /* ML:1005-1007 */
            c = (float)(0.75 * (double)C[p] + 0.25 * (double)cn);
            if (c > 1) c = 1.0;
            if (c < 0) c = 0.0;
        }
        Hn[p] = hn;
        Vn[p] = vn;
        Cn[p] = c;
    }

    free(Rw);
    free(U);
    free(D);
    free(P);
    free(t1);
    free(N);
    for (int t = 0; t < 5; t++) free(A[t]);
}

/* MGL:2257-2293 (S passes, all three planes from the pre-pass snapshot) then
 * MGL:2361-2412 (box filter rows then columns, clamp). */
void ugsmo_smooth(float* H, float* V, float* C, int w, int h, int passes) {
    size_t n = (size_t)w * h;
    float g[5], a[5];
    ugsmo_taps(g, a);
    float* sH = (float*)malloc(n * sizeof(float));
    float* sV = (float*)malloc(n * sizeof(float));
    float* sC = (float*)malloc(n * sizeof(float));
    for (int j = 0; j < passes; j++) {
        memcpy(sH, H, n * sizeof(float));
        memcpy(sV, V, n * sizeof(float));
        memcpy(sC, C, n * sizeof(float));
        ugsmo_smooth_pass(sH, sC, H, w, h);
        ugsmo_smooth_pass(sV, sC, V, w, h);
        ugsmo_smooth_pass(sC, sC, C, w, h);
    }
    ugsmo_conv_rows_clamp(H, sH, w, h, a);
    ugsmo_conv_cols_clamp(sH, H, w, h, a);
    ugsmo_conv_rows_clamp(V, sH, w, h, a);
    ugsmo_conv_cols_clamp(sH, V, w, h, a);
    ugsmo_conv_rows_clamp(C, sH, w, h, a);
    ugsmo_conv_cols_clamp(sH, C, w, h, a);
    free(sH);
    free(sV);
    free(sC);
}

/* MGL:1662-2489 */
void ugsmo_match_level(const float* L3, const float* R3, float* H, float* V, float* C, int w, int h, int iters,
                       int passes, int coarsest) {
    size_t n = (size_t)w * h;
    float* Hn = (float*)malloc(n * sizeof(float));
    float* Vn = (float*)malloc(n * sizeof(float));
    float* Cn = (float*)malloc(n * sizeof(float));
    float T = 1.0f; /* MGL:1673 */
    for (int m = 1; m <= iters; m++) {
        ugsmo_correlate(L3, R3, H, V, C, w, h, T, coarsest && m == 1, Hn, Vn, Cn, NULL, NULL, NULL, NULL);
        ugsmo_smooth(Hn, Vn, Cn, w, h, passes);
        T = ugsmo_threshold_after(iters, m, T);
        memcpy(H, Hn, n * sizeof(float));
        memcpy(V, Vn, n * sizeof(float));
        memcpy(C, Cn, n * sizeof(float));
    }
    free(Hn);
    free(Vn);
    free(Cn);
}

/* MGL:1526-1590; with a crop it is MGL:1595-1655 (upsample to fine dims, centre-crop). */
void ugsmo_upsample_state(const float* H, const float* V, const float* C, int w, int h, int fine_w, int fine_h,
                          int crop_x, int crop_y, int out_w, int out_h, float* Ho, float* Vo, float* Co) {
    size_t nf = (size_t)fine_w * fine_h;
    float* tmp = (float*)malloc(nf * sizeof(float));
    const float* in[3] = {H, V, C};
    float* out[3] = {Ho, Vo, Co};
    for (int k = 0; k < 3; k++) {
        ugsmo_upsample_disp(in[k], w, h, tmp, fine_w, fine_h);
        for (int i = 0; i < out_h; i++)
            memcpy(out[k] + (size_t)i * out_w, tmp + (size_t)(crop_y + i) * fine_w + crop_x, out_w * sizeof(float));
    }
    free(tmp);
}

typedef struct {
    float* lv[UGSMO_MAX_LEVELS]; /* 3 planes per level */
    int w[UGSMO_MAX_LEVELS], h[UGSMO_MAX_LEVELS];
} pyr_t;

static void build_pyramid(const uint8_t* rgb, int rows, int cols, size_t step, int levels, pyr_t* p) {
    ugsmo_level_dims(cols, rows, levels, p->w, p->h);
    size_t n0 = (size_t)rows * cols;
    float* planes = (float*)malloc(3 * n0 * sizeof(float));
    ugsmo_unpack_rgb8(rgb, rows, cols, step, planes);
    for (int l = 0; l < levels; l++) p->lv[l] = (float*)malloc(3 * (size_t)p->w[l] * p->h[l] * sizeof(float));
    for (int ch = 0; ch < 3; ch++) {
        float* out[UGSMO_MAX_LEVELS];
        for (int l = 0; l < levels; l++) out[l] = p->lv[l] + (size_t)ch * p->w[l] * p->h[l];
        ugsmo_pyramid_plane(planes + ch * n0, cols, rows, levels, out);
    }
    free(planes);
}
static void free_pyramid(pyr_t* p, int levels) {
    for (int l = 0; l < levels; l++) free(p->lv[l]);
}

/* MGL:303-403 with fov == 0, matching() at MGL:1196-1318. */
int ugsmo_match_full(const ugsmo_config* cfg, const uint8_t* rgbL, const uint8_t* rgbR, int rows, int cols,
                     size_t step, float* H, float* V, float* C) {
    int levels = cfg->levels;
    if (levels < 1 || levels > UGSMO_MAX_LEVELS || rows < 1 || cols < 1) return 1;
    pyr_t pl, pr;
    build_pyramid(rgbL, rows, cols, step, levels, &pl);
    build_pyramid(rgbR, rows, cols, step, levels, &pr);
    if (pl.w[levels - 1] < 1 || pl.h[levels - 1] < 1) {
        free_pyramid(&pl, levels);
        free_pyramid(&pr, levels);
        return 1;
    }
    float *cH = NULL, *cV = NULL, *cC = NULL;
    for (int i = levels - 1; i >= 0; i--) {
        size_t n = (size_t)pl.w[i] * pl.h[i];
        float *lH, *lV, *lC;
        if (i == levels - 1) { /* zero start (reference: uninitialised, MGL:1247) */
            lH = (float*)calloc(n, sizeof(float));
            lV = (float*)calloc(n, sizeof(float));
            lC = (float*)calloc(n, sizeof(float));
        } else {
            lH = (float*)malloc(n * sizeof(float));
            lV = (float*)malloc(n * sizeof(float));
            lC = (float*)malloc(n * sizeof(float));
            ugsmo_upsample_state(cH, cV, cC, pl.w[i + 1], pl.h[i + 1], pl.w[i], pl.h[i], 0, 0, pl.w[i], pl.h[i], lH,
                                 lV, lC);
            free(cH);
            free(cV);
            free(cC);
        }
        ugsmo_match_level(pl.lv[i], pr.lv[i], lH, lV, lC, pl.w[i], pl.h[i], ugsmo_level_iters(cfg, i),
                          ugsmo_level_smooth(cfg, i), i == levels - 1);
        cH = lH;
        cV = lV;
        cC = lC;
    }
    size_t n0 = (size_t)rows * cols;
    memcpy(H, cH, n0 * sizeof(float));
    memcpy(V, cV, n0 * sizeof(float));
    memcpy(C, cC, n0 * sizeof(float));
    free(cH);
    free(cV);
    free(cC);
    free_pyramid(&pl, levels);
    free_pyramid(&pr, levels);
    return 0;
}

/* Centre crop of one level, MGL:1171-1185: l = x/2 - fovW/2, u = y/2 - fovH/2. */
static float* crop3(const float* src3, int w, int h, int fw, int fh) {
    int l = w / 2 - fw / 2, u = h / 2 - fh / 2;
    float* out = (float*)malloc(3 * (size_t)fw * fh * sizeof(float));
    for (int k = 0; k < 3; k++)
        for (int i = 0; i < fh; i++)
            memcpy(out + (size_t)k * fw * fh + (size_t)i * fw, src3 + (size_t)k * w * h + (size_t)(u + i) * w + l,
                   fw * sizeof(float));
    return out;
}

/* MGL:534-700: full pyramids, CreateFoveatedPyramid (MGL:1128-1190), matching() with the
 * foveated dim override (MGL:1230-1240) and upsample choice (MGL:1275-1294); output in the
 * level-major stack layout of UG_GPU_matcher.cpp:293-303. */
int ugsmo_match_foveated(const ugsmo_config* cfg, const uint8_t* rgbL, const uint8_t* rgbR, int rows, int cols,
                         size_t step, float* stackH, float* stackV, float* stackC, int* fovW, int* fovH) {
    int levels = cfg->levels, fl = cfg->fovea_level;
    if (levels < 1 || levels > UGSMO_MAX_LEVELS || fl < 2 || fl > levels || rows < 1 || cols < 1) return 1;
    pyr_t pl, pr;
    build_pyramid(rgbL, rows, cols, step, levels, &pl);
    build_pyramid(rgbR, rows, cols, step, levels, &pr);
    int fw = pl.w[fl - 1], fh = pl.h[fl - 1];
    if (fovW) *fovW = fw;
    if (fovH) *fovH = fh;
    int w2 = pl.w[fl - 2], h2 = pl.h[fl - 2];
    float *cH = NULL, *cV = NULL, *cC = NULL;
    int cw = 0, chh = 0;
    for (int i = levels - 1; i >= 0; i--) {
        int mw = (i < fl - 1) ? fw : pl.w[i];
        int mh = (i < fl - 1) ? fh : pl.h[i];
        size_t n = (size_t)mw * mh;
        float *lH, *lV, *lC;
        if (i == levels - 1) {
            lH = (float*)calloc(n, sizeof(float));
            lV = (float*)calloc(n, sizeof(float));
            lC = (float*)calloc(n, sizeof(float));
        } else {
            lH = (float*)malloc(n * sizeof(float));
            lV = (float*)malloc(n * sizeof(float));
            lC = (float*)malloc(n * sizeof(float));
            if (i + 1 >= fl) /* MGL:1283-1286 */
                ugsmo_upsample_state(cH, cV, cC, cw, chh, mw, mh, 0, 0, mw, mh, lH, lV, lC);
            else /* MGL:1290 + 1612-1615 */
                ugsmo_upsample_state(cH, cV, cC, cw, chh, w2, h2, w2 / 2 - cw / 2, h2 / 2 - chh / 2, mw, mh, lH, lV,
                                     lC);
            free(cH);
            free(cV);
            free(cC);
        }
        float *L3, *R3;
        int cropped = (i < fl - 1);
        if (cropped) {
            L3 = crop3(pl.lv[i], pl.w[i], pl.h[i], fw, fh);
            R3 = crop3(pr.lv[i], pr.w[i], pr.h[i], fw, fh);
        } else {
            L3 = pl.lv[i];
            R3 = pr.lv[i];
        }
        ugsmo_match_level(L3, R3, lH, lV, lC, mw, mh, ugsmo_level_iters(cfg, i), ugsmo_level_smooth(cfg, i),
                          i == levels - 1);
        if (cropped) {
            free(L3);
            free(R3);
        }
        if (i < fl) {
            size_t off = (size_t)i * fw * fh;
            memcpy(stackH + off, lH, n * sizeof(float));
            memcpy(stackV + off, lV, n * sizeof(float));
            memcpy(stackC + off, lC, n * sizeof(float));
        }
        cH = lH;
        cV = lV;
        cC = lC;
        cw = mw;
        chh = mh;
    }
    free(cH);
    free(cV);
    free(cC);
    free_pyramid(&pl, levels);
    free_pyramid(&pr, levels);
    return 0;
}

/* ML:435-464: tex2D(src, x / scalefactor, y / scalefactor) * scalefactor, float division and multiply. */
void ugsmo_partsubsample_disp(const float* src, int w, int h, float scale, float* dst, int w2, int h2) {
#pragma omp parallel for schedule(static)
    for (int iy = 0; iy < h2; iy++) {
        float y = (float)iy + 0.5f;
        int sy = tex_index(y / scale, h);
        for (int ix = 0; ix < w2; ix++) {
            float x = (float)ix + 0.5f;
            int sx = tex_index(x / scale, w);
            dst[(size_t)iy * w2 + ix] = scale * src[(size_t)sy * w + sx];
        }
    }
}

/* MGL:2589-2701.  Start from the (uncropped) level fl-1; for level = fl-1 .. 1: upsample the composite to
 * level-1 dims with levelscale = (float)SCALE (MGL:2638, 2666), then paste the fovea of level-1 at
 * (l1, u) = (w/2 - fovW/2, h/2 - fovH/2) (MGL:2643-2647, 2678-2683).  H, V and C are treated alike. */
void ugsmo_composite(const float* stackH, const float* stackV, const float* stackC, int cols, int rows, int levels,
                     int fovea_level, float* H, float* V, float* C) {
    int w[UGSMO_MAX_LEVELS], h[UGSMO_MAX_LEVELS];
    ugsmo_level_dims(cols, rows, levels, w, h);
    const int fl = fovea_level, fw = w[fl - 1], fh = h[fl - 1];
    const float* stack[3] = {stackH, stackV, stackC};
    float* out[3] = {H, V, C};
    const float levelscale = SCALE_D;
    for (int k = 0; k < 3; k++) {
        size_t n = (size_t)fw * fh;
        float* cur = (float*)malloc(n * sizeof(float));
        memcpy(cur, stack[k] + (size_t)(fl - 1) * n, n * sizeof(float));
        int cw = fw, chh = fh;
        for (int level = fl - 1; level > 0; level--) {
            int nw = w[level - 1], nh = h[level - 1];
            float* nxt = (level == 1) ? out[k] : (float*)malloc((size_t)nw * nh * sizeof(float));
            ugsmo_partsubsample_disp(cur, cw, chh, levelscale, nxt, nw, nh);
            int l1 = nw / 2 - fw / 2, u = nh / 2 - fh / 2;
            for (int i = 0; i < fh; i++)
                memcpy(nxt + (size_t)(u + i) * nw + l1, stack[k] + (size_t)(level - 1) * n + (size_t)i * fw,
                       fw * sizeof(float));
            free(cur);
            cur = nxt;
            cw = nw;
            chh = nh;
        }
        if (fl == 1) {
            memcpy(out[k], cur, n * sizeof(float));
            free(cur);
        }
    }
}

int ugsmo_match_composite(const ugsmo_config* cfg, const uint8_t* rgbL, const uint8_t* rgbR, int rows, int cols,
                          size_t step, float* H, float* V, float* C) {
    int w[UGSMO_MAX_LEVELS], h[UGSMO_MAX_LEVELS];
    if (cfg->levels < 1 || cfg->levels > UGSMO_MAX_LEVELS || cfg->fovea_level < 2 || cfg->fovea_level > cfg->levels)
        return 1;
    ugsmo_level_dims(cols, rows, cfg->levels, w, h);
    int fl = cfg->fovea_level, fw = w[fl - 1], fh = h[fl - 1];
    size_t n = (size_t)fl * fw * fh;
    float* sH = (float*)malloc(n * sizeof(float));
    float* sV = (float*)malloc(n * sizeof(float));
    float* sC = (float*)malloc(n * sizeof(float));
    int rc = ugsmo_match_foveated(cfg, rgbL, rgbR, rows, cols, step, sH, sV, sC, NULL, NULL);
    if (rc == 0) ugsmo_composite(sH, sV, sC, cols, rows, cfg->levels, fl, H, V, C);
    free(sH);
    free(sV);
    free(sC);
    return rc;
}
