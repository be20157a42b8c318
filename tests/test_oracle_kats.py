"""CPU tests (-m "not gpu"): pin the oracle.

The reference ships no tests or golden vectors, so the pins are:
  * the reference's own CPU convolution (convolutionSeparable_gold.cpp) compiled in place
    into oracle/_ref/ -- border semantics and tap order of the zero-border filter;
  * known-answer tests derived by hand from the kernel arithmetic in MatchLib.cu;
  * the level/iteration/threshold tables of SURVEY.md 8a (derived from MatchGPULib.cpp);
  * golden outputs of the reference CUDA build (tests/test_refcuda_golden.py).
"""
import ctypes as C
import os

import numpy as np
import pytest

from ug_stereomatcher_b200 import synth

REF_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")


def test_taps(oracle):
    g, a = oracle.taps()
    # MatchGPULib.cpp:761-775: literals renormalised by their float sum, index order
    lit = np.array([0.0816475, 0.218507, 0.303281, 0.218507, 0.0816475], np.float32)
    s = np.float32(0)
    for v in lit:
        s = np.float32(s + v)
    assert (g == (lit / s).astype(np.float32)).all()
    assert (a == np.array([0, 0.3333, 0.3333, 0.3333, 0], np.float32)).all()
    assert g[0] == g[4] and g[1] == g[3]


def test_level_tables(oracle):
    """SURVEY.md 8a table (MatchGPULib.cpp:1224-1228, 1741, 2257-2261)."""
    w, h = oracle.level_dims(4928, 3264)
    assert w == [4928, 3484, 2463, 1741, 1231, 870, 615, 434, 306, 216, 152, 107, 75, 53]
    assert h == [3264, 2307, 1631, 1153, 815, 576, 407, 287, 202, 142, 100, 70, 49, 34]
    w, h = oracle.level_dims(640, 480)
    assert w == [640, 452, 319, 225, 159, 112, 79, 55, 38, 26, 18, 12, 8, 5]
    assert h == [480, 339, 239, 168, 118, 83, 58, 41, 28, 19, 13, 9, 6, 4]
    w, h = oracle.level_dims(2464, 1632)
    assert w[:7] == [2464, 1742, 1231, 870, 615, 434, 306] and h[:7] == [1632, 1153, 815, 576, 407, 287, 202]
    cfg = oracle.default_config()
    assert [oracle.level_iters(cfg, i) for i in range(14)] == [2, 4, 6, 8, 10, 12] + [22] * 8
    assert [oracle.level_smooth(cfg, i) for i in range(14)] == [10, 10] + [5] * 12
    w, h = oracle.level_dims(4928, 3264)
    assert sum(a * b * oracle.level_iters(cfg, i) for i, (a, b) in enumerate(zip(w, h))) == 131429636


def test_threshold_schedule(oracle):
    """MatchGPULib.cpp:2299-2306: T per iteration m for mi = 22, 12, 4, 2."""
    def seq(mi):
        T, out = 1.0, []
        for m in range(1, mi + 1):
            out.append(T)
            T = oracle.threshold_after(mi, m, T)
        return out
    f = lambda v: float(np.float32(v))
    assert seq(22) == [1.0] * 10 + [f(0.55)] * 2 + [f(0.46)] * 2 + [f(0.37)] * 2 + [f(0.28)] * 2 + [f(0.19)] * 2 + [f(0.1)] * 2
    assert seq(12) == [1.0] * 2 + [f(0.82)] * 2 + [f(0.64)] * 2 + [f(0.46)] * 2 + [f(0.28)] * 2 + [f(0.1)] * 2
    assert seq(4) == [1.0, 1.0, f(0.1), f(0.1)]
    assert seq(2) == [1.0, 1.0]


def _gold(name):
    path = os.path.join(REF_DIR, name)
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/%s not built (reference checkout absent)" % name)
    return C.CDLL(path)


@pytest.mark.parametrize("shape", [(37, 53), (5, 4), (64, 128), (3, 300)])
def test_zero_border_conv_matches_reference_cpu(oracle, shape):
    """oracle conv == the reference's convolutionRowCPU/ColumnCPU (convolutionSeparable_gold.cpp:20-75).

    libgold_fma.so is that file compiled with one FMA per tap (what nvcc does to the GPU kernels):
    must match bit for bit.  libgold_strict.so (separate multiply and add) pins the border/tap
    semantics independently of contraction: equal within 1 ulp of the accumulated sum."""
    rng = np.random.default_rng(1)
    h, w = shape
    src = rng.uniform(0, 255, (h, w)).astype(np.float32)
    g, _ = oracle.taps()
    for libname, exact in (("libgold_fma.so", True), ("libgold_strict.so", False)):
        lib = _gold(libname)
        for mine, fn in ((oracle.conv_rows_zero, lib.convolutionRowCPU), (oracle.conv_cols_zero, lib.convolutionColumnCPU)):
            ref = np.empty_like(src)
            fn(ref.ctypes.data_as(C.c_void_p), src.ctypes.data_as(C.c_void_p), g.ctypes.data_as(C.c_void_p), w, h, 2)
            got = mine(src, g)
            if exact:
                assert (got == ref).all()
            else:
                np.testing.assert_allclose(got, ref, rtol=3e-7, atol=0)


def test_conv_impulse_responses(oracle):
    g, a = oracle.taps()
    img = np.zeros((9, 11), np.float32)
    img[4, 5] = 1.0
    r = oracle.conv_rows_zero(img, g)
    assert (r[4, 3:8] == g[::-1]).all() and r.sum() == np.float32(g.sum())  # taps[2-j] at x+j
    c = oracle.conv_cols_clamp(img, g)
    assert (c[2:7, 5] == g[::-1]).all()
    # clamp border: an impulse in the corner is seen by taps -2,-1,0 of pixel 0
    img[:] = 0
    img[0, 0] = 1.0
    r = oracle.conv_rows_clamp(img, g)
    want = np.float32(0)
    for k in (-2, -1, 0):
        want = np.float32(np.float32(g[2 - k]) + want) if False else want
    acc = np.float32(0)
    for k in (-2, -1, 0):  # fmaf(1, tap, acc) is exact add here
        acc = np.float32(acc + g[2 - k])
    assert r[0, 0] == acc and r[0, 1] == np.float32(g[4] + g[3]) and r[0, 2] == g[4]
    # zero border: no contribution from outside
    r0 = oracle.conv_rows_zero(img, g)
    assert r0[0, 0] == g[2] and r0[0, 1] == g[3] and r0[0, 2] == g[4]
    # box taps: gain 0.9999 per pass in the interior (MatchGPULib.cpp:344-348)
    ones = np.ones((6, 6), np.float32)
    b = oracle.conv_cols_clamp(oracle.conv_rows_clamp(ones, a), a)
    assert abs(float(b[3, 3]) - 0.9999 ** 2) < 1e-6


def test_decimate_indices(oracle):
    """MatchLib.cu:320-332: factor 2.0 takes the odd sample (2x+1, 2y+1); sqrt2 takes floor((i+.5)*s)."""
    src = np.arange(20 * 31, dtype=np.float32).reshape(20, 31)
    d = oracle.decimate(src, 2.0, 15, 10)
    assert (d == src[1::2, 1::2][:10, :15]).all()
    s = np.float32(1.41421356)
    w2, h2 = int(31 / 1.41421356), int(20 / 1.41421356)
    d = oracle.decimate(src, float(s), w2, h2)
    xs = np.floor((np.arange(w2, dtype=np.float32) + np.float32(0.5)) * s).astype(int)
    ys = np.floor((np.arange(h2, dtype=np.float32) + np.float32(0.5)) * s).astype(int)
    assert (d == src[np.ix_(ys, xs)]).all()


def test_upsample_disp(oracle):
    """MatchLib.cu:383-394: nearest upsample with s=(float)(1/SCALE), value*SCALE in double, all planes."""
    src = np.arange(12, dtype=np.float32).reshape(3, 4) + 0.25
    d = oracle.upsample_disp(src, 5, 4)
    s = np.float32(1 / 1.41421356)
    xs = np.minimum(np.floor((np.arange(5, dtype=np.float32) + np.float32(0.5)) * s).astype(int), 3)
    ys = np.minimum(np.floor((np.arange(4, dtype=np.float32) + np.float32(0.5)) * s).astype(int), 2)
    want = (1.41421356 * src[np.ix_(ys, xs)].astype(np.float64)).astype(np.float32)
    assert (d == want).all()


def test_warp_nearest_and_clamp(oracle):
    """MatchLib.cu:507-514: Rw = R[floor(y+.5+dV)][floor(x+.5+dH)], clamped."""
    R = np.arange(6 * 8, dtype=np.float32).reshape(6, 8)
    H = np.zeros((6, 8), np.float32)
    V = np.zeros((6, 8), np.float32)
    H[2, 3] = 0.49     # floor(3.5+0.49)=3
    H[2, 4] = 0.5      # floor(4.5+0.5)=5
    H[2, 5] = -0.51    # floor(5.5-0.51)=4
    H[0, 0] = -7.0     # clamps to 0
    H[5, 7] = 9.0      # clamps to 7
    V[5, 7] = 3.0      # clamps to 5
    V[3, 1] = -1.0     # row 2
    out, sx, sy = oracle.warp(R, H, V)
    assert (sx[2, 3], sx[2, 4], sx[2, 5], sx[0, 0], sx[5, 7]) == (3, 5, 4, 0, 7)
    assert (sy[5, 7], sy[3, 1]) == (5, 2)
    assert out[3, 1] == R[2, 1] and out[2, 4] == R[2, 5]


def test_poly_known_answers(oracle):
    """PolyDisparity (MatchLib.cu:809-836) by hand."""
    # symmetric peak: b1 = 0, c1 = r - c < 0 -> h = 0, c* = c, conf = 0.3c+0.7
    off, conf, br = oracle.poly(0.5, 0.9, 0.5, 1.0)
    assert off == 0.0 and br == 1 and conf == np.float32(0.3 * np.float64(np.float32(0.9)) + 0.7)
    # flat / convex: c1 >= 0 -> (0, 0.4)
    assert oracle.poly(0.5, 0.5, 0.5, 1.0) == (0.0, float(np.float32(0.4)), 0)
    assert oracle.poly(0.9, 0.2, 0.9, 1.0)[2] == 0
    # NaN input (0/0 correlation, MatchLib.cu:684) falls in the flat branch
    assert oracle.poly(float("nan"), 0.5, 0.5, 1.0) == (0.0, float(np.float32(0.4)), 0)
    # asymmetric: l=.2 c=.8 r=.6 -> b1=.2 c1=-.4 h=.25
    l, c, r = np.float32(0.2), np.float32(0.8), np.float32(0.6)
    b1 = np.float32((r - l) / np.float32(2))
    c1 = np.float32(r - np.float32(c + b1))
    h = np.float32((-np.float64(b1) * 0.5) / np.float64(c1))
    off, conf, br = oracle.poly(float(l), float(c), float(r), 1.0)
    assert off == h and br == 1 and abs(off - 0.25) < 1e-6
    # threshold clamp
    assert oracle.poly(float(l), float(c), float(r), 0.1)[0] == float(np.float32(0.1))
    # c* > 1 -> rescaled offset, conf = 1
    off, conf, br = oracle.poly(0.2, 0.99, 0.9, 1.0)
    assert br == 2 and conf == 1.0 and 0 < off < 1


def test_smooth_pass_known_answer(oracle):
    """smoothKernel (MatchLib.cu:1106-1133): row 0 / col 0 untouched, clamp right/bottom, fixed order."""
    rng = np.random.default_rng(3)
    src = rng.uniform(-2, 2, (5, 6)).astype(np.float32)
    conf = rng.uniform(0.2, 1, (5, 6)).astype(np.float32)
    out = oracle.smooth_pass(src, conf)
    assert (out[0] == src[0]).all() and (out[:, 0] == src[:, 0]).all()
    def one(y, x):
        import math
        sd, sc = np.float32(0), np.float32(0)
        for (yy, xx) in ((y, x), (y, x - 1), (y, min(x + 1, 5)), (y - 1, x), (min(y + 1, 4), x)):
            sd = np.float32(math.fma(float(src[yy, xx]), float(conf[yy, xx]), float(sd))) if hasattr(math, "fma") else None
            sc = np.float32(sc + conf[yy, xx])
        return sd, sc
    import math
    if hasattr(math, "fma"):
        for (y, x) in ((2, 3), (4, 5), (1, 1), (4, 1)):
            sd, sc = one(y, x)
            assert out[y, x] == np.float32(sd / sc)
    # constant field is a fixed point
    k = np.full((5, 6), 1.5, np.float32)
    assert np.allclose(oracle.smooth_pass(k, conf), 1.5, atol=2e-7)


def test_pyramid_structure(oracle):
    """Level 1 from blurred level 0 by the sqrt2 NN step; level i+2 from blurred level i by taking the odd
    sample (MatchGPULib.cpp:1082-1096): two interleaved octave chains."""
    L, _, _, _ = synth.make_pair(96, 64, 0)
    pl = oracle.unpack_rgb8(L)[0]
    pyr = oracle.pyramid_plane(pl, 6)
    g, _ = oracle.taps()
    w, h = oracle.level_dims(96, 64, 6)
    assert [p.shape for p in pyr] == [(h[i], w[i]) for i in range(6)]
    assert (pyr[0] == pl).all()
    b0 = oracle.conv_cols_zero(oracle.conv_rows_zero(pyr[0], g), g)
    assert (pyr[2] == b0[1::2, 1::2][:h[2], :w[2]]).all()
    b1 = oracle.conv_cols_zero(oracle.conv_rows_zero(pyr[1], g), g)
    assert (pyr[3] == b1[1::2, 1::2][:h[3], :w[3]]).all()
    assert (pyr[1] == oracle.decimate(b0, float(np.float32(1.41421356)), w[1], h[1])).all()


def test_unpack_step(oracle):
    img = np.arange(4 * 5 * 3, dtype=np.uint8).reshape(4, 5, 3)
    pad = np.zeros((4, 9, 3), np.uint8)
    pad[:, :5] = img
    view = pad[:, :5]  # non-contiguous rows (step = 27 bytes)
    import ugsm_oracle as o
    out = np.empty((3, 4, 5), np.float32)
    o.lib().ugsmo_unpack_rgb8(C.c_void_p(pad.ctypes.data), 4, 5, C.c_size_t(pad.strides[0]), C.c_void_p(out.ctypes.data))
    assert (out == img.transpose(2, 0, 1)).all()


def test_identity_and_shift(oracle):
    """Known-answer behaviour of the whole matcher: identity -> ~0, integer shift -> shift (sign: right = left + d)."""
    L, _, _, _ = synth.make_pair(160, 120, 7)
    cfg = oracle.default_config(levels=3)
    H, V, Cc = oracle.match_full(L, L, cfg)
    m = (slice(20, -20), slice(20, -20))
    assert np.median(np.abs(H[m])) < 0.02 and np.median(np.abs(V[m])) < 0.02
    R = np.roll(L, 2, axis=1)
    H, V, Cc = oracle.match_full(L, R, cfg)
    assert abs(np.median(H[m]) - 2.0) < 0.25 and abs(np.median(V[m])) < 0.1


def test_foveated_crop_offsets(oracle):
    """16 MP fovea 615x407, crop offsets l = w/2 - fovW/2, u = h/2 - fovH/2."""
    w, h = oracle.level_dims(4928, 3264)
    fw, fh = w[6], h[6]
    assert (fw, fh) == (615, 407)
    offs = [(w[i] // 2 - fw // 2, h[i] // 2 - fh // 2) for i in range(6)]
    assert offs == [(2157, 1429), (1435, 950), (924, 612), (563, 373), (308, 204), (128, 85)]


def test_foveated_stack_layout(oracle):
    L, R, _, _ = synth.make_pair(320, 240, 2)
    cfg = oracle.default_config(levels=5, fovea_level=3, max_iters=2)
    sH, sV, sC, fw, fh = oracle.match_foveated(L, R, cfg)
    w, h = oracle.level_dims(320, 240, 5)
    assert (fw, fh) == (w[2], h[2]) and sH.shape == (3 * fh, fw)
    # level fovea_level-1 is the uncropped natural level: same result as a full-pyramid run stopped there
    assert np.isfinite(sH).all() and np.isfinite(sC).all()
