// ugsm_k2_stream.cuh -- K2 for the large levels as a row pipeline over the smoothing passes (sm_100a).
//
// Same arithmetic as k2_smooth_box (ugsm_kernels.cuh; replaces the 3*S smoothKernel launches, ML:1092-1145 /
// MGL:2257-2293, and the 6 box-filter launches, ML:1593-1697 / MGL:2361-2412, of one iteration).  The tiled kernel
// holds a 128 x 72 region in 216 KB of shared memory (one CTA per SM), recomputes a halo of S+1 cells on all four
// sides (1.77x the useful work at S = 10) and pays one CTA-wide barrier per pass with nothing else resident.
//
// Here a CTA owns a 128-column strip and a row segment, and the passes form a producer/consumer pipeline down the
// rows: warp 0 streams input rows into shared memory (2-D TMA tiles of 128 x 1 floats per plane completing on the
// mbarrier of the row's ring slot, zero fill outside the allocation; cp.async where no tensor map can be built),
// warp t (1..S) turns rows r-1, r, r+1 of pass t-1 into row r of pass t, and the last warp applies the 3x3 box
// filter and stores.  A pass lags its producer by two rows and all stages advance together, one row per tick, with
// one CTA barrier per tick (a variant coupled through full / empty mbarriers per ring row instead was measured
// slower: its try_wait round trips sit on the critical path of every row).  A consumer only ever reads the three
// rows behind the one its producer is writing, so the ring between two passes is 4 rows (6 KB); the loader's ring is
// 8 rows because it runs three rows ahead.  72 KB at S = 10, 42 KB at S = 5: several CTAs fit an SM and their ticks
// interleave.  The halo is recomputed in x only (128 / (128 - 2 hpad)) and once per segment in y.
//
// The kernel is bound by the math dispatch port (tools/ubench/pipes.cu: every ALU-pipe instruction -- FSEL, FMNMX,
// IADD3, ISETP -- occupies it for two cycles, an FFMA for one), so the row loop of the common case is straight-line
// code with no selects, no per-row conditions and three address additions: FAST steps run on rows 1 .. H-2 of strips
// whose 128 columns have both neighbours inside the image.  A lane owns 4 adjacent columns and keeps rows r-1 and r
// in registers from one row to the next (the loop is unrolled three times so that the roles rotate without moves):
// 9 shared loads and 3 shared stores per 4 cells and pass.  The edge strips run the same loop with the column rules
// of ML:1106-1131 (selects) and are cut into more row segments than interior strips so that their slower steps do
// not set the kernel's time; row 0 and row H-1 take the GENERIC step, which reloads its three rows and applies
// every border rule.
#pragma once
#include "ugsm_k1_stream.cuh"  // TMA / mbarrier helpers (k1s_mbar_*, k1s_tma_load_2d)
#include "ugsm_kernels.cuh"

namespace ugsm {

constexpr int K2S_MAX_PASSES = 10;
constexpr int K2S_D0 = 8;                      // rows of the loader's ring
constexpr int K2S_D = 4;                       // rows of a ring between two passes
constexpr int K2S_SLOTF = 3 * K2_BW;           // floats of a ring slot: one row of H, V, C (plane stride K2_BW)
constexpr int K2S_RING0F = K2S_D0 * K2S_SLOTF; // floats of the loader's ring
constexpr int K2S_RINGF = K2S_D * K2S_SLOTF;   // floats of a pass ring

struct K2StreamParams {
    K2Params k;       // th is unused; hpad = column halo (multiple of 4, >= passes + 1)
    int nseg;         // row segments of an interior strip over [y_begin, y_end); edge strips use nseg + edge_extra
    int edge_extra;   // grid.y = nseg + edge_extra
    int use_tma;      // the loader fetches rows as TMA tiles (tm[]) instead of cp.async
    // TMA descriptors of the three input planes: pitch x h floats, box 128 x 1, zero fill outside (same cells as the
    // cp.async loader's zero fill: columns outside the allocation, rows are never requested outside the image)
    alignas(64) CUtensorMap tm[3];
};

__host__ __device__ inline size_t k2s_smem_bytes(int passes) {
    return ((size_t)K2S_RING0F + (size_t)passes * K2S_RINGF) * sizeof(float) + K2S_D0 * sizeof(unsigned long long);
}

__device__ __forceinline__ void k2s_cp_async16(float* smem_dst, const float* gsrc, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 16 : 0;  // src-size 0: zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}

// Rows [lo, hi] of a pass live in a ring of `mask + 1` slots; a slot holds the row's H, V and C cells side by side
// (plane stride 128 floats: compile-time offsets), row r sits in slot r & mask.  own / lf / rt: the lane's first
// cell, its left and its right neighbour in slot 0.
struct K2sRing {
    float* own;
    const float* lf;
    const float* rt;
    int mask, lo, hi;
    __device__ __forceinline__ int slot(int r) const { return (r & mask) * K2S_SLOTF; }
};

// A lane's 4 cells of one ring row plus the cell on either side, all three planes.
__device__ __forceinline__ void k2s_load_row(const K2sRing& g, int r, K2Row& w) {
    const int so = g.slot(r);
    const float* own = g.own + so;
    const float* lf = g.lf + so;
    const float* rt = g.rt + so;
    float4 t = *reinterpret_cast<const float4*>(own);
    w.h[1] = t.x, w.h[2] = t.y, w.h[3] = t.z, w.h[4] = t.w;
    w.h[0] = lf[0], w.h[5] = rt[0];
    t = *reinterpret_cast<const float4*>(own + K2_BW);
    w.v[1] = t.x, w.v[2] = t.y, w.v[3] = t.z, w.v[4] = t.w;
    w.v[0] = lf[K2_BW], w.v[5] = rt[K2_BW];
    t = *reinterpret_cast<const float4*>(own + 2 * K2_BW);
    w.c[1] = t.x, w.c[2] = t.y, w.c[3] = t.z, w.c[4] = t.w;
    w.c[0] = lf[2 * K2_BW], w.c[5] = rt[2 * K2_BW];
}
__device__ __forceinline__ void k2s_store_row(const K2sRing& g, int r, const float oh[4], const float ov[4],
                                              const float oc[4]) {
    float* d = g.own + g.slot(r);
    *reinterpret_cast<float4*>(d) = make_float4(oh[0], oh[1], oh[2], oh[3]);
    *reinterpret_cast<float4*>(d + K2_BW) = make_float4(ov[0], ov[1], ov[2], ov[3]);
    *reinterpret_cast<float4*>(d + 2 * K2_BW) = make_float4(oc[0], oc[1], oc[2], oc[3]);
}

// The pipeline advances in ticks separated by one CTA barrier: at tick tau the loader has row Rin0 + tau complete,
// pass t produces row Rin0 + tau - 2 t (its three input rows were completed at earlier ticks) and the last warp
// finishes output row Rin0 + tau - 2 S - 1 - m.  While pass t writes row a, pass t + 1 reads rows a-3 .. a-1 and the
// box filter rows a-3 .. a-1 of the last pass: 4 slots never collide; the loader runs K2S_LEAD rows ahead of the row
// it completes and pass 1 reads up to 3 rows behind that one: 8 slots.
constexpr int K2S_LEAD = 3;

constexpr int K2S_NT_SMALL = 256;  // S <= 6
constexpr int K2S_NT_LARGE = 384;  // S <= 10
#ifndef K2S_MINB_SMALL
#define K2S_MINB_SMALL 2
#endif
#ifndef K2S_MINB_LARGE
#define K2S_MINB_LARGE 2
#endif

// Rows r .. fb of one smoothing pass with rows r-1 and r carried in registers (the three row registers rotate with
// the rows: the loop is unrolled three times).  One barrier per row.  EDGE: the strip touches the left / right image
// border, so the column rules of ML:1106-1131 apply and lanes wholly outside the image idle.
template <bool EDGE>
__device__ __forceinline__ void k2s_fast_rows(const K2sRing& in, const K2sRing& out, int& r, int& tau, const int fb,
                                              const int gx0, const int W, const int Hh, const bool lane_in) {
    K2Row w0, w1, w2;
    if (!EDGE || lane_in) {
        k2s_load_row(in, r - 1, w0);
        k2s_load_row(in, r, w1);
    }
    auto fast = [&](const K2Row& up, const K2Row& ce, K2Row& dn) {
        if (!EDGE || lane_in) {
            k2s_load_row(in, r + 1, dn);
            float oh[4], ov[4], oc[4];
            k2_smooth_cells<EDGE, false>(up, ce, dn, gx0, r, W, Hh, oh, ov, oc);
            k2s_store_row(out, r, oh, ov, oc);
        }
        r++, tau++;
        __syncthreads();
    };
    while (r + 2 <= fb) {
        fast(w0, w1, w2);
        fast(w1, w2, w0);
        fast(w2, w0, w1);
    }
    if (r <= fb) {
        fast(w0, w1, w2);
        if (r <= fb) fast(w1, w2, w0);
    }
}

// NT = threads the launch may use ((passes + 2) * 32 <= NT); MINB = CTAs per SM the register budget is cut for.
template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k2_stream(const __grid_constant__ K2StreamParams sp) {
    extern __shared__ __align__(128) unsigned char k2s_smem[];
    const K2Params& p = sp.k;
    const int S = p.passes, m = 1;  // m: rows / columns the 3x3 box filter reaches beyond the last pass
    const int W = p.w, Hh = p.h, pitch = p.pitch;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* const rings = reinterpret_cast<float*>(k2s_smem);  // ring 0: [8][3][128]; rings 1..S: [4][3][128]
    const int TWu = K2_BW - 2 * p.hpad;        // useful columns per strip
    const int X0 = blockIdx.x * TWu - p.hpad;  // image column of strip column 0
    // strips whose 128 columns all have both neighbours inside the image need no column rule at all
    const bool strip_interior = (X0 >= 1) && (X0 + K2_BW <= W - 1);
    const int nseg = sp.nseg + (strip_interior ? 0 : sp.edge_extra);
    if ((int)blockIdx.y >= nseg) return;  // CTA-uniform
    const int brows = p.y_end - p.y_begin;
    const int seg_rows = (brows + nseg - 1) / nseg;
    const int Ys = p.y_begin + blockIdx.y * seg_rows;
    const int Ye = min(Ys + seg_rows, p.y_end);
    if (Ys >= Ye) return;  // CTA-uniform
    const int gx0 = X0 + 4 * lane;
    const int Rin0 = Ys - m - S;                // first input row of the pipeline
    const int T = (Ye - Rin0) + 2 * S + 1 + m;  // ticks
    // rows pass s must produce (s = 0: the input rows), clipped to the image
    auto ring = [&](int s) {
        K2sRing r;
        float* base = (s == 0) ? rings : rings + K2S_RING0F + (size_t)(s - 1) * K2S_RINGF;
        r.own = base + 4 * lane;
        r.lf = r.own + ((lane == 0) ? 0 : -1);   // strip column -1 does not exist: never used by a useful cell
        r.rt = r.own + ((lane == 31) ? 3 : 4);
        r.mask = (s == 0) ? K2S_D0 - 1 : K2S_D - 1;
        r.lo = max(0, Ys - m - (S - s));
        r.hi = min(Hh - 1, Ye - 1 + m + (S - s));
        return r;
    };

    if (warp == 0) {
        // ---- loader: input rows -> ring 0, zero fill outside the allocation ----
        const K2sRing out = ring(0);
        if (sp.use_tma) {
            // a row of the three planes = three 2-D TMA tiles (128 x 1 floats each, 512-byte slots of the ring) that
            // complete on the mbarrier of the row's ring slot; one elected lane issues, the whole warp waits
            unsigned long long* bars = reinterpret_cast<unsigned long long*>(rings + K2S_RING0F + (size_t)S * K2S_RINGF);
            if (lane == 0) {
                for (int i = 0; i < K2S_D0; i++) k1s_mbar_init((unsigned)__cvta_generic_to_shared(&bars[i]), 1);
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            __syncwarp();
            auto issue = [&](int r) {
                if (r >= out.lo && r <= out.hi && lane == 0) {
                    const int sl = r & out.mask;
                    const unsigned bar = (unsigned)__cvta_generic_to_shared(&bars[sl]);
                    float* d = rings + sl * K2S_SLOTF;
                    // the slot's previous tenant was last read before the CTA barrier of the previous tick
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    k1s_mbar_expect_tx(bar, 3 * K2_BW * (int)sizeof(float));
#pragma unroll
                    for (int k = 0; k < 3; k++) k1s_tma_load_2d(d + k * K2_BW, &sp.tm[k], X0, r, bar);
                }
            };
            for (int r = Rin0; r < Rin0 + K2S_LEAD; r++) issue(r);
            for (int tau = 0; tau < T; tau++) {
                issue(Rin0 + tau + K2S_LEAD);
                const int r = Rin0 + tau;  // this row must have landed
                if (r >= out.lo && r <= out.hi)
                    k1s_mbar_wait((unsigned)__cvta_generic_to_shared(&bars[r & out.mask]), ((r - out.lo) >> 3) & 1);
                __syncthreads();
            }
        } else {
            // 16 bytes per lane per plane with cp.async (planes a tensor map cannot describe)
            const bool col_in = (gx0 >= 0 && gx0 + 3 < pitch);
            auto issue = [&](int r) {
                if (r >= out.lo && r <= out.hi) {
                    const size_t o = col_in ? (size_t)r * pitch + gx0 : 0;
                    float* d = out.own + out.slot(r);
                    k2s_cp_async16(d, p.Hi + o, col_in);
                    k2s_cp_async16(d + K2_BW, p.Vi + o, col_in);
                    k2s_cp_async16(d + 2 * K2_BW, p.Ci + o, col_in);
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            };
            for (int r = Rin0; r < Rin0 + K2S_LEAD; r++) issue(r);
            for (int tau = 0; tau < T; tau++) {
                issue(Rin0 + tau + K2S_LEAD);
                asm volatile("cp.async.wait_group %0;" ::"n"(K2S_LEAD) : "memory");  // row Rin0 + tau has landed
                __syncthreads();
            }
        }
    } else if (warp <= S) {
        // ---- smoothing pass t = warp: rows r-1, r, r+1 of ring t-1 -> row r of ring t (ML:1106-1133) ----
        const K2sRing in = ring(warp - 1), out = ring(warp);
        const bool lane_in = (gx0 < W) && (gx0 + 3 >= 0);  // lanes wholly outside the image idle
        int r = Rin0 - 2 * warp, tau = 0;
        // GENERIC step: self-contained (reloads its rows), every border rule
        auto generic = [&]() {
            if (r >= out.lo && r <= out.hi && lane_in) {
                K2Row up, ce, dn;
                k2s_load_row(in, r, ce);
                if (r - 1 >= in.lo) k2s_load_row(in, r - 1, up);  // r == 0: that row keeps its values
                else up = ce;
                if (r + 1 <= in.hi) k2s_load_row(in, r + 1, dn);  // r == Hh-1: clamped to ce inside
                else dn = ce;
                float oh[4], ov[4], oc[4];
                k2_smooth_cells<true>(up, ce, dn, gx0, r, W, Hh, oh, ov, oc);
                k2s_store_row(out, r, oh, ov, oc);
            }
            r++, tau++;
            __syncthreads();
        };
        // FAST rows: every row neighbour exists and is inside this CTA's rings (edge strips keep the column rules)
        const int fa = max(out.lo, max(in.lo + 1, 1));
        const int fb = min(out.hi, min(in.hi - 1, Hh - 2));
        while (tau < T && r < fa) generic();
        if (tau < T && r <= fb) {
            if (strip_interior) k2s_fast_rows<false>(in, out, r, tau, fb, gx0, W, Hh, true);
            else k2s_fast_rows<true>(in, out, r, tau, fb, gx0, W, Hh, lane_in);
        }
        while (tau < T) generic();
    } else if (warp == S + 1) {
        // ---- box filter (ML:1616-1623 rows, ML:1683-1690 columns; taps {0,a,a,a,0}, clamp) + store ----
        const K2sRing in = ring(S);
        const float a1 = p.a1;
        const bool lane_useful = (4 * lane >= p.hpad) && (4 * lane < K2_BW - p.hpad) && (gx0 < W);
        float* const outs[3] = {p.Ho, p.Vo, p.Co};
        int tau = 0, y = Rin0 - 2 * S - 1 - m;
        // row taps of one ring row for the lane's 4 cells, all three planes; BORDER: clamp at the image's columns
        auto rowconv = [&](int row, bool border, float a[3][4]) {
            K2Row rw;
            k2s_load_row(in, row, rw);
            const float* pl[3] = {rw.h, rw.v, rw.c};
#pragma unroll
            for (int k = 0; k < 3; k++)
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const int gx = gx0 + c;
                    const bool hl = !border || (gx - 1 >= 0), hr = !border || (gx + 1 <= W - 1);
                    float t = __fmaf_rn(hl ? pl[k][c] : pl[k][c + 1], a1, 0.0f);
                    t = __fmaf_rn(pl[k][c + 1], a1, t);
                    a[k][c] = __fmaf_rn(hr ? pl[k][c + 2] : pl[k][c + 1], a1, t);
                }
        };
        auto emit = [&](const float au[3][4], const float ac[3][4], const float ad[3][4]) {
#pragma unroll
            for (int k = 0; k < 3; k++) {
                float res[4];
#pragma unroll
                for (int c = 0; c < 4; c++)  // column taps in the order up, centre, down
                    res[c] = __fmaf_rn(ad[k][c], a1, __fmaf_rn(ac[k][c], a1, __fmaf_rn(au[k][c], a1, 0.0f)));
                float* dst = outs[k] + (size_t)y * pitch + gx0;
                if (gx0 + 3 < W) {
                    *reinterpret_cast<float4*>(dst) = make_float4(res[0], res[1], res[2], res[3]);
                } else {
                    dst[0] = res[0];
                    if (gx0 + 1 < W) dst[1] = res[1];
                    if (gx0 + 2 < W) dst[2] = res[2];
                }
            }
        };
        auto generic = [&]() {
            if (y >= Ys && y < Ye && lane_useful) {
                const int yu = (y - 1 >= 0) ? y - 1 : y, yd = (y + 1 <= Hh - 1) ? y + 1 : y;
                float au[3][4], ac[3][4], ad[3][4];
                rowconv(yu, true, au);
                rowconv(y, true, ac);
                rowconv(yd, true, ad);
                emit(au, ac, ad);
            }
            y++, tau++;
            __syncthreads();
        };
        const int fa = max(Ys, 1), fb = min(Ye - 1, Hh - 2);
        while (tau < T && y < fa) generic();
        if (tau < T && y <= fb) {
            const bool border = !strip_interior;  // CTA-uniform: column clamps only in the edge strips
            float a0[3][4], a1r[3][4], a2[3][4];
            if (lane_useful) {
                rowconv(y - 1, border, a0);
                rowconv(y, border, a1r);
            }
            auto fast = [&](const float au[3][4], const float ac[3][4], float ad[3][4]) {
                if (lane_useful) {
                    rowconv(y + 1, border, ad);
                    emit(au, ac, ad);
                }
                y++, tau++;
                __syncthreads();
            };
            while (y + 2 <= fb) {
                fast(a0, a1r, a2);
                fast(a1r, a2, a0);
                fast(a2, a0, a1r);
            }
            if (y <= fb) {
                fast(a0, a1r, a2);
                if (y <= fb) fast(a1r, a2, a0);
            }
        }
        while (tau < T) generic();
    }
}

}  // namespace ugsm
