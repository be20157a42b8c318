/*
 * ugsm_oracle_cloud.c -- CPU oracle for the point-cloud step that consumes the matcher's H, V
 * planes.  TEST INFRASTRUCTURE ONLY (see ugsm_oracle.h).
 *
 * Reference file restated here:  PC = src/pointcloud/getPointCloud.cpp
 *   PC:886-945   get3DPoint: closed-form two-view triangulation (mixed float / double)
 *   PC:375-421   mapXcoord / mapYcoord: fovea-stack coordinate -> full-resolution coordinate
 *   PC:432-484   upper_marginOf_in / left_marginOf_in
 *   PC:697-740   doReconstructionRGB_FOV  (foveated cloud, column-major point order)
 *   PC:741-782   doReconstructionRGB      (full cloud, column-major point order)
 * The host build of the reference is `-O3` for baseline x86-64 (CMakeLists.txt:86): no FMA
 * contraction, FLT_EVAL_METHOD 0.  This file is built with -ffp-contract=off and spells no fma().
 *
 * Pinned: oracle/_ref/refpc is the UNMODIFIED reference file compiled through
 * oracle/refbuild/stubs_pc; tests/test_cloud_oracle.py compares the two bit for bit on random and
 * adversarial inputs, and tests/golden/refpc_cloud.npz holds outputs of that binary.
 *
 * C++ typing rules the restatement has to follow (every variable a..j, x, y is float):
 *   - pow(v, 2.0) with a float v promotes to double; the square of a float is exact in double,
 *     so it is written (double)v * (double)v;
 *   - float (op) float stays float and rounds to float; float (op) double is double;
 *   - a + b + c groups as (a + b) + c, a*b*c as (a*b)*c;
 *   - the four results are rounded to float when assigned (PC:931-943), the three quotients are
 *     float divisions (PC:945-947).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "ugsm_oracle.h"

static inline double sq(float v) { return (double)v * (double)v; }

void ugsmo_triangulate(const double P1[12], const double P2[12], float x1, float y1, float x2, float y2,
                       float out[3]) {
    /* PC:918-929.  P(r,c) = P[4*r + c] (PC:1158-1159). */
    const float a = (float)P1[0];
    const float b = (float)(P1[2] - x1);
    const float c = (float)P1[5];
    const float d = (float)(P1[6] - y1);
    const float e = (float)(P2[0] - x2 * P2[8]);
    const float f = (float)(P2[1] - x2 * P2[9]);
    const float g = (float)(P2[2] - x2 * P2[10]);
    const float h = (float)(P2[4] - y2 * P2[8]);
    const float i = (float)(P2[5] - y2 * P2[9]);
    const float j = (float)(P2[6] - y2 * P2[10]);
    const float x = (float)(x2 * P2[11] - P2[3]);
    const float y = (float)(y2 * P2[11] - P2[7]);

    const double a2 = sq(a), b2 = sq(b), c2 = sq(c), d2 = sq(d);
    const float ab = a * b, cd = c * d;

    /* float sub-expressions that recur (each is evaluated the same way at every occurrence) */
    const float fh_ei = f * h - e * i;
    const float gi_fj = g * i - f * j;
    const float k1 = d * f * h - c * g * h - d * e * i + c * e * j; /* PC:931, 934, 940 */
    const float hx_ey = h * x - e * y;
    const float ex_hy = e * x + h * y;
    const float fx_iy = f * x + i * y;
    const float gx_jy = g * x + j * y;
    const float nix_fy = -(i * x) + f * y;
    const float ix_fy = i * x - f * y;
    const float ngi_fj = -(g * i) + f * j;

    /* PC:931-933 */
    const float xa = k1 * (-(d * i * x) + c * j * x + d * f * y - c * g * y);
    const double xb = b2 * (fh_ei * nix_fy + c2 * ex_hy);
    const double xc = ab * (ngi_fj * ix_fy + cd * fx_iy - c2 * gx_jy);
    const float XUp = (float)(xa + xb + xc);

    /* PC:934-936 */
    const double ya = (b2 * fh_ei + d * k1) * hx_ey;
    const double yb = ab * ((c * d * e + g * h * i - 2.0 * f * h * j + e * i * j) * x +
                            (c * d * h + f * g * h - 2.0 * e * g * i + e * f * j) * y);
    const double yc = a2 * (gi_fj * (-(j * x) + g * y) + d2 * fx_iy - cd * gx_jy);
    const float YUp = (float)(ya + yb + yc);

    /* PC:937-939 */
    const float za = c * (-(d * f * h) + c * g * h + d * e * i - c * e * j) * hx_ey;
    const double zb = ab * (fh_ei * nix_fy + c2 * ex_hy);
    const double zc = a2 * (gi_fj * ix_fy - cd * fx_iy + c2 * gx_jy);
    const float ZUp = (float)(za - zb + zc);

    /* PC:940-944 */
    const double da = b2 * (c2 * (sq(e) + sq(h)) + sq(fh_ei));
    const double db = sq(k1);
    const double dc = 2.0 * a * b * (-(cd * (e * f + h * i)) + fh_ei * ngi_fj + c2 * (e * g + h * j));
    const double dd = a2 * (d2 * (sq(f) + sq(i)) + sq(gi_fj) - 2.0 * c * d * (f * g + i * j) + c2 * (sq(g) + sq(j)));
    const float divisor = (float)(da + db - dc + dd);

    out[0] = XUp / divisor;
    out[1] = YUp / divisor;
    out[2] = ZUp / divisor;
}

int64_t ugsmo_cloud_count(int rows, int cols, int sampling) {
    if (rows <= 0 || cols <= 0 || sampling <= 0) return 0;
    return (int64_t)((cols + sampling - 1) / sampling) * ((rows + sampling - 1) / sampling);
}

static inline uint32_t pack_rgb(const uint8_t* px, int bgr) {
    /* PC:765-771: the node's image is BGR8; rgb = r << 16 | g << 8 | b */
    const uint32_t c0 = px[0], c1 = px[1], c2 = px[2];
    return bgr ? (c2 << 16 | c1 << 8 | c0) : (c0 << 16 | c1 << 8 | c2);
}

/* PC:741-782.  Points are emitted column by column (ii outer, jj inner). */
int64_t ugsmo_cloud_full(const double P1[12], const double P2[12], const float* dX, const float* dY, int rows,
                         int cols, const uint8_t* img, size_t step, int bgr, int sampling, float* points,
                         int stride) {
    if (rows <= 0 || cols <= 0 || sampling <= 0 || stride < 4) return -1;
    const int ncol = (cols + sampling - 1) / sampling, nrow = (rows + sampling - 1) / sampling;
#pragma omp parallel for schedule(static)
    for (int ci = 0; ci < ncol; ci++) {
        const int ii = ci * sampling;
        for (int rj = 0; rj < nrow; rj++) {
            const int jj = rj * sampling;
            const size_t o = (size_t)jj * cols + ii;
            float p[3];
            /* PC:909-912: x1 = xx, x2 = xx + dispX (int + float -> float) */
            ugsmo_triangulate(P1, P2, (float)ii, (float)jj, (float)ii + dX[o], (float)jj + dY[o], p);
            float* rec = points + ((size_t)ci * nrow + rj) * stride;
            const uint32_t rgb = pack_rgb(img + (size_t)jj * step + 3 * (size_t)ii, bgr);
            rec[0] = p[0], rec[1] = p[1], rec[2] = p[2];
            if (stride >= 8) { /* pcl::PointXYZRGB memory layout: x y z 1 | rgb 0 0 0 */
                rec[3] = 1.0f;
                memcpy(rec + 4, &rgb, 4);
                rec[5] = rec[6] = rec[7] = 0.0f;
            } else {
                memcpy(rec + 3, &rgb, 4);
            }
        }
    }
    return (int64_t)ncol * nrow;
}

/* PC:432-484 with destLevel = 0 (PC:569): margin of the fovea of stack level src_level inside the
 * full-resolution image; the dims chain is int <- int / 1.41421356 (double). */
void ugsmo_cloud_margins(int im_rows, int im_cols, int fovea_level, int src_level, int* left, int* upper) {
    int w = im_cols, h = im_rows;
    const int lev = (fovea_level - 1) - src_level; /* PC:436, 467: "6 - srcLevel" for foveaLevel 7 */
    for (int k = 0; k < lev; k++) {
        w = (int)(w / 1.41421356);
        h = (int)(h / 1.41421356);
    }
    *left = im_cols / 2 - w / 2;
    *upper = im_rows / 2 - h / 2;
}

/* PC:385-391: the per-level scale is computed in float: pow(float, float) is powf in C++11. */
float ugsmo_cloud_scale(int src_level) {
    const float root = (float)1.41421356237309504880; /* PC:372; destLevel 0 <= srcLevel -> M_SQRT2 */
    return powf(root, (float)abs(src_level - 0));
}

/* PC:697-740 + PC:886-908.  stackX / stackY: (levels * fov_h) x fov_w, level-major. */
int64_t ugsmo_cloud_foveated(const double P1[12], const double P2[12], const float* stackX, const float* stackY,
                             int fov_h, int fov_w, int fovea_level, int src_level, int im_rows, int im_cols,
                             const uint8_t* img, size_t step, int bgr, int sampling, float* points, int stride) {
    if (fov_h <= 0 || fov_w <= 0 || sampling <= 0 || stride < 4 || src_level < 0 || src_level >= fovea_level)
        return -1;
    int left, upper;
    ugsmo_cloud_margins(im_rows, im_cols, fovea_level, src_level, &left, &upper);
    const float sc = ugsmo_cloud_scale(src_level);
    const int ncol = (fov_w + sampling - 1) / sampling, nrow = (fov_h + sampling - 1) / sampling;
#pragma omp parallel for schedule(static)
    for (int ci = 0; ci < ncol; ci++) {
        const int ii = ci * sampling;
        for (int rj = 0; rj < nrow; rj++) {
            const int jj = rj * sampling;
            const size_t o = (size_t)(jj + fov_h * src_level) * fov_w + ii;
            /* PC:388: original = (float)margin + (float)src * scale, all float */
            const float x1 = (float)left + (float)ii * sc;
            const float y1 = (float)upper + (float)jj * sc;
            /* PC:900-901: the float sum is truncated to int by mapXcoord's int parameter */
            const int sx = (int)((float)ii + stackX[o]);
            const int sy = (int)((float)jj + stackY[o]);
            const float x2 = (float)left + (float)sx * sc;
            const float y2 = (float)upper + (float)sy * sc;
            float p[3];
            ugsmo_triangulate(P1, P2, x1, y1, x2, y2, p);
            const int px = (int)x1, py = (int)y1; /* PC:722-723 */
            float* rec = points + ((size_t)ci * nrow + rj) * stride;
            const uint32_t rgb = pack_rgb(img + (size_t)py * step + 3 * (size_t)px, bgr);
            rec[0] = p[0], rec[1] = p[1], rec[2] = p[2];
            if (stride >= 8) {
                rec[3] = 1.0f;
                memcpy(rec + 4, &rgb, 4);
                rec[5] = rec[6] = rec[7] = 0.0f;
            } else {
                memcpy(rec + 3, &rgb, 4);
            }
        }
    }
    return (int64_t)ncol * nrow;
}

/* Range maps: getRangePoint / getRangePointFOV (PC:951-1019) at every pixel, row-major as range_map.at<float>(jj, ii)
 * (PC:749, 832).  Their ZUp and divisor expressions are the ones of get3DPoint, token for token (PC:937-944 vs
 * PC:972-982), so the range is the z coordinate ugsmo_triangulate returns. */
int ugsmo_range_full(const double P1[12], const double P2[12], const float* dX, const float* dY, int rows, int cols,
                     float* range) {
    if (rows <= 0 || cols <= 0) return -1;
#pragma omp parallel for schedule(static)
    for (int jj = 0; jj < rows; jj++)
        for (int ii = 0; ii < cols; ii++) {
            const size_t o = (size_t)jj * cols + ii;
            float p[3];
            ugsmo_triangulate(P1, P2, (float)ii, (float)jj, (float)ii + dX[o], (float)jj + dY[o], p);
            range[o] = p[2];
        }
    return 0;
}

int ugsmo_range_foveated(const double P1[12], const double P2[12], const float* stackX, const float* stackY, int fov_h,
                         int fov_w, int fovea_level, int src_level, int im_rows, int im_cols, float* range) {
    if (fov_h <= 0 || fov_w <= 0 || src_level < 0 || src_level >= fovea_level) return -1;
    int left, upper;
    ugsmo_cloud_margins(im_rows, im_cols, fovea_level, src_level, &left, &upper);
    const float sc = ugsmo_cloud_scale(src_level);
#pragma omp parallel for schedule(static)
    for (int jj = 0; jj < fov_h; jj++)
        for (int ii = 0; ii < fov_w; ii++) {
            const size_t o = (size_t)(jj + fov_h * src_level) * fov_w + ii;
            const float x1 = (float)left + (float)ii * sc, y1 = (float)upper + (float)jj * sc; /* PC:988-992 */
            const int sx = (int)((float)ii + stackX[o]), sy = (int)((float)jj + stackY[o]);       /* PC:994-995 */
            const float x2 = (float)left + (float)sx * sc, y2 = (float)upper + (float)sy * sc;
            float p[3];
            ugsmo_triangulate(P1, P2, x1, y1, x2, y2, p);
            range[(size_t)jj * fov_w + ii] = p[2];
        }
    return 0;
}
