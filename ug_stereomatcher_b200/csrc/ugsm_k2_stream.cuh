// ugsm_k2_stream.cuh -- K2 for the large levels as a row pipeline over the smoothing passes (sm_100a).
//
// Same arithmetic as k2_smooth_box (ugsm_kernels.cuh; replaces the 3*S smoothKernel launches, ML:1092-1145 /
// MGL:2257-2293, and the 6 box-filter launches, ML:1593-1697 / MGL:2361-2412, of one iteration).  The tiled kernel
// holds a 128 x 72 region in 216 KB of shared memory (one CTA per SM), recomputes a halo of S+1 cells on all four
// sides (1.77x the useful work at S = 10) and pays one CTA-wide barrier per pass with nothing else resident.
//
// Here a CTA owns a 128-column strip and a row segment, and the passes form a producer/consumer pipeline down the
// rows: warp 0 streams input rows into shared memory (cp.async, 16 bytes per lane, zero fill outside the image),
// warp t (1..S) turns rows r-1, r, r+1 of pass t-1 into row r of pass t, and the last warp applies the 3x3 box
// filter and stores.  Each stage writes an
// 8-row ring (12 KB); a pass lags its producer by two rows and all stages advance together, one row per tick, with
// one CTA barrier per tick (a variant coupled through full / empty mbarriers per ring row instead was measured
// slower: its try_wait round trips sit on the critical path of every row).  Several CTAs fit an SM (72 KB at S = 5),
// so their ticks interleave.  Used for S <= 6 on levels of 1.5 MP and more (-15..-28 % against the tiled kernel); at
// S = 10 the 12-warp CTA is alone on its SM and the tiled kernel stays faster (tuning log).
// The halo is recomputed in x only (128 / (128 - 2 hpad)) and once per segment in y.  A lane owns 4 adjacent
// columns and keeps rows r-1 and r in registers from one row to the next (the row loop is unrolled three times so
// that the roles rotate without moves): one row is read and one written per cell and pass, with immediate offsets.
#pragma once
#include "ugsm_kernels.cuh"

namespace ugsm {

constexpr int K2S_MAX_PASSES = 10;
constexpr int K2S_D = 8;                       // rows per ring
constexpr int K2S_PL = K2S_D * K2_BW;          // plane stride inside a ring (floats)
constexpr int K2S_RINGF = 3 * K2S_PL;          // floats per ring (H, V, C planes)

struct K2StreamParams {
    K2Params k;       // th is unused; hpad = column halo (multiple of 4, >= passes + 1)
    int seg_rows;     // rows per segment (grid.y segments cover [y_begin, y_end))
};

__host__ __device__ inline size_t k2s_smem_bytes(int passes) {
    return (size_t)(passes + 1) * K2S_RINGF * sizeof(float);
}

__device__ __forceinline__ void k2s_cp_async16(float* smem_dst, const float* gsrc, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 16 : 0;  // src-size 0: zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
// A lane's 4 cells of one ring row plus the cell on either side, all three planes (immediate offsets).
__device__ __forceinline__ void k2s_load_row(const float* rp, int dl, int dr, K2Row& r) {
    float4 t = *reinterpret_cast<const float4*>(rp);
    r.h[1] = t.x, r.h[2] = t.y, r.h[3] = t.z, r.h[4] = t.w;
    r.h[0] = rp[dl], r.h[5] = rp[dr];
    t = *reinterpret_cast<const float4*>(rp + K2S_PL);
    r.v[1] = t.x, r.v[2] = t.y, r.v[3] = t.z, r.v[4] = t.w;
    r.v[0] = rp[K2S_PL + dl], r.v[5] = rp[K2S_PL + dr];
    t = *reinterpret_cast<const float4*>(rp + 2 * K2S_PL);
    r.c[1] = t.x, r.c[2] = t.y, r.c[3] = t.z, r.c[4] = t.w;
    r.c[0] = rp[2 * K2S_PL + dl], r.c[5] = rp[2 * K2S_PL + dr];
}

// Rows [lo, hi] of a pass live in ring `base` (8 slots of 128 columns per plane, plane stride K2S_PL);
// row r sits in slot r & 7.
struct K2sRing {
    float* base;
    int lo, hi;
    __device__ __forceinline__ const float* row(int r, int lane) const { return base + (r & (K2S_D - 1)) * K2_BW + 4 * lane; }
};

// The pipeline advances in ticks separated by one CTA barrier: at tick tau the loader has row Rin0 + tau complete,
// pass t produces row Rin0 + tau - 2 t (its three input rows were completed at earlier ticks) and the last warp
// finishes output row Rin0 + tau - 2 S - 1 - m.  Rows a consumer still reads are at most 3 behind the row its
// producer writes, the loader runs K2S_LEAD rows ahead: 8 slots per ring never collide.
constexpr int K2S_LEAD = 3;

__global__ void __launch_bounds__((K2S_MAX_PASSES + 2) * 32) k2_stream(const __grid_constant__ K2StreamParams sp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const K2Params& p = sp.k;
    const int S = p.passes, m = 1;  // m: rows / columns the 3x3 box filter reaches beyond the last pass
    const int W = p.w, Hh = p.h, pitch = p.pitch;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* const rings = reinterpret_cast<float*>(smem_raw);  // [S + 1][3][8][128]
    const int TWu = K2_BW - 2 * p.hpad;        // useful columns per strip
    const int X0 = blockIdx.x * TWu - p.hpad;  // image column of strip column 0
    const int Ys = p.y_begin + blockIdx.y * sp.seg_rows;
    const int Ye = min(Ys + sp.seg_rows, p.y_end);
    if (Ys >= Ye) return;  // CTA-uniform
    const int gx0 = X0 + 4 * lane;
    const int Rin0 = Ys - m - S;                // first input row of the pipeline
    const int T = (Ye - Rin0) + 2 * S + 1 + m;  // ticks
    // rows pass s must produce (s = 0: the input rows), clipped to the image
    auto ring = [&](int s) {
        K2sRing r;
        r.base = rings + (size_t)s * K2S_RINGF;
        r.lo = max(0, Ys - m - (S - s));
        r.hi = min(Hh - 1, Ye - 1 + m + (S - s));
        return r;
    };
    // strips whose 128 columns all have both neighbours inside the image skip the border rules on interior rows
    const bool strip_interior = (X0 >= 1) && (X0 + K2_BW <= W - 1);
    const int dl = (lane == 0) ? 0 : -1, dr = (lane == 31) ? 3 : 4;  // neighbours of the lane's 4 cells in a ring row

    if (warp == 0) {
        // ---- loader: input rows -> ring 0, 16 bytes per lane per plane, zero fill outside the allocation ----
        const K2sRing out = ring(0);
        const bool col_in = (gx0 >= 0 && gx0 + 3 < pitch);
        auto issue = [&](int r) {
            if (r >= out.lo && r <= out.hi) {
                const size_t o = col_in ? (size_t)r * pitch + gx0 : 0;
                float* d = const_cast<float*>(out.row(r, lane));
                k2s_cp_async16(d, p.Hi + o, col_in);
                k2s_cp_async16(d + K2S_PL, p.Vi + o, col_in);
                k2s_cp_async16(d + 2 * K2S_PL, p.Ci + o, col_in);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        for (int r = Rin0; r < Rin0 + K2S_LEAD; r++) issue(r);
        for (int tau = 0; tau < T; tau++) {
            issue(Rin0 + tau + K2S_LEAD);
            asm volatile("cp.async.wait_group %0;" ::"n"(K2S_LEAD) : "memory");  // row Rin0 + tau has landed
            __syncthreads();
        }
    } else if (warp <= S) {
        // ---- smoothing pass t = warp: rows r-1, r, r+1 of ring t-1 -> row r of ring t (ML:1106-1133) ----
        const K2sRing in = ring(warp - 1), out = ring(warp);
        const bool lane_in = (gx0 < W) && (gx0 + 3 >= 0);  // lanes wholly outside the image idle
        K2Row w0, w1, w2;
        int r = Rin0 - 2 * warp;
        auto step = [&](K2Row& up, K2Row& ce, K2Row& dn) {
            if (r >= out.lo && r <= out.hi && lane_in) {
                if (r == out.lo) {
                    if (r - 1 >= in.lo) k2s_load_row(in.row(r - 1, lane), dl, dr, up);  // r == 0: that row keeps its values
                    k2s_load_row(in.row(r, lane), dl, dr, ce);
                }
                if (r + 1 <= in.hi) k2s_load_row(in.row(r + 1, lane), dl, dr, dn);  // r == Hh-1: clamped to ce inside
                float oh[4], ov[4], oc[4];
                if (strip_interior && r >= 1 && r + 1 <= Hh - 1)
                    k2_smooth_cells<false>(up, ce, dn, gx0, r, W, Hh, oh, ov, oc);
                else
                    k2_smooth_cells<true>(up, ce, dn, gx0, r, W, Hh, oh, ov, oc);
                float* d = const_cast<float*>(out.row(r, lane));
                *reinterpret_cast<float4*>(d) = make_float4(oh[0], oh[1], oh[2], oh[3]);
                *reinterpret_cast<float4*>(d + K2S_PL) = make_float4(ov[0], ov[1], ov[2], ov[3]);
                *reinterpret_cast<float4*>(d + 2 * K2S_PL) = make_float4(oc[0], oc[1], oc[2], oc[3]);
            }
            r++;
            __syncthreads();
        };
        // The three row registers rotate with the rows.  The rotation must be in phase with the first active row:
        // idle ticks before it are burnt one at a time.
        int tau = 0;
        for (; tau < T && r < out.lo; tau++) {
            r++;
            __syncthreads();
        }
        for (; tau + 3 <= T; tau += 3) {
            step(w0, w1, w2);
            step(w1, w2, w0);
            step(w2, w0, w1);
        }
        if (tau < T) {
            step(w0, w1, w2);
            tau++;
        }
        if (tau < T) step(w1, w2, w0);
    } else if (warp == S + 1) {
        // ---- box filter (ML:1616-1623 rows, ML:1683-1690 columns; taps {0,a,a,a,0}, clamp) + store, or plain store ----
        const K2sRing in = ring(S);
        const float a1 = p.a1;
        const bool lane_useful = (4 * lane >= p.hpad) && (4 * lane < K2_BW - p.hpad) && (gx0 < W);
        float* const outs[3] = {p.Ho, p.Vo, p.Co};
        for (int tau = 0; tau < T; tau++) {
            const int y = Rin0 + tau - 2 * S - 1 - m;
            if (y >= Ys && y < Ye && lane_useful) {
                float res[3][4];
                {
                    const int yu = (y - 1 >= 0) ? y - 1 : y, yd = (y + 1 <= Hh - 1) ? y + 1 : y;
                    const int rows[3] = {yu, y, yd};
                    const bool inner = strip_interior;  // every cell has a left and a right neighbour in the image
#pragma unroll
                    for (int k = 0; k < 3; k++)
#pragma unroll
                        for (int c = 0; c < 4; c++) res[k][c] = 0.0f;
#pragma unroll
                    for (int j = 0; j < 3; j++) {
                        K2Row rw;
                        k2s_load_row(in.row(rows[j], lane), dl, dr, rw);
                        const float* pl[3] = {rw.h, rw.v, rw.c};
#pragma unroll
                        for (int k = 0; k < 3; k++)
#pragma unroll
                            for (int c = 0; c < 4; c++) {
                                const int gx = gx0 + c;
                                const bool hl = inner || (gx - 1 >= 0), hr = inner || (gx + 1 <= W - 1);
                                float a = __fmaf_rn(hl ? pl[k][c] : pl[k][c + 1], a1, 0.0f);
                                a = __fmaf_rn(pl[k][c + 1], a1, a);
                                a = __fmaf_rn(hr ? pl[k][c + 2] : pl[k][c + 1], a1, a);
                                res[k][c] = __fmaf_rn(a, a1, res[k][c]);  // column taps in the order up, centre, down
                            }
                    }
                }
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    float* dst = outs[k] + (size_t)y * pitch + gx0;
                    if (gx0 + 3 < W) {
                        *reinterpret_cast<float4*>(dst) = make_float4(res[k][0], res[k][1], res[k][2], res[k][3]);
                    } else {
                        dst[0] = res[k][0];
                        if (gx0 + 1 < W) dst[1] = res[k][1];
                        if (gx0 + 2 < W) dst[2] = res[k][2];
                    }
                }
            }
            __syncthreads();
        }
    }
}

}  // namespace ugsm
