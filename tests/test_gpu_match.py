"""-m gpu: whole-pair parity through the C-ABI against the oracle (BASELINE.json configs[0] and
reduced versions of the other configs), plus size-independent properties.

Tolerances from BASELINE.json north_star: sub-pixel H/V within 1e-3 px max-abs of the oracle;
integer decisions (warp source index, PolyDisparity branch) bit-exact.  The kernels are written
to be bit-identical to the oracle, so the tests first assert exact equality and report the
north-star tolerance numbers on failure.
"""
import numpy as np
import pytest

from ug_stereomatcher_b200 import synth

pytestmark = pytest.mark.gpu

TOL_PX = 1e-3  # north_star: sub-pixel H/V disparities within 1e-3 px max-abs of the oracle


def _cmp(got, want, what):
    for g, w, n in zip(got, want, ["H", "V", "C"]):
        assert g.shape == w.shape
        d = np.abs(g.astype(np.float64) - w.astype(np.float64))
        assert np.isfinite(g).all(), "%s %s: non-finite output" % (what, n)
        assert d.max() <= TOL_PX, "%s %s: max|gpu-oracle| = %g px > %g" % (what, n, d.max(), TOL_PX)
        assert (g == w).all(), "%s %s: within tolerance but not bit-exact (%d px differ, max %g)" % (
            what, n, int((g != w).sum()), d.max())


def test_config0_640x480_5_levels(capi, oracle):
    """BASELINE configs[0]: 640x480 random-dot-like RGB pair with a known H/V field, 5-level pyramid."""
    L, R, dH, dV = synth.make_pair(640, 480, 0)
    with capi.Handle(levels=5) as h:
        got = h.match_full(L, R)
    want = oracle.match_full(L, R, oracle.default_config(levels=5))
    _cmp(got, want, "640x480/5")
    m = (slice(24, -24), slice(24, -24))
    assert np.abs(got[0] - dH)[m].mean() < 0.35 and np.abs(got[1] - dV)[m].mean() < 0.35  # sanity EPE only


def test_640x480_14_levels(capi, oracle):
    """Reference level count (MAX_LEVEL=14): coarsest level is 5x4 pixels."""
    L, R, _, _ = synth.make_pair(640, 480, 1)
    with capi.Handle() as h:
        got = h.match_full(L, R)
    want = oracle.match_full(L, R)
    _cmp(got, want, "640x480/14")


@pytest.mark.parametrize("shape", [(97, 131), (64, 64), (301, 203)])
def test_odd_shapes(capi, oracle, shape):
    rows, cols = shape
    L, R, _, _ = synth.make_pair(cols, rows, 4)
    with capi.Handle(levels=4) as h:
        got = h.match_full(L, R)
    want = oracle.match_full(L, R, oracle.default_config(levels=4))
    _cmp(got, want, "%dx%d/4" % (cols, rows))


def test_row_step_padding(capi, oracle):
    """cv_bridge images carry a row step (MatchGPULib.cpp:318,336): padded rows must not matter."""
    L, R, _, _ = synth.make_pair(200, 120, 5)
    Lp = np.zeros((120, 232, 3), np.uint8)
    Rp = np.zeros((120, 232, 3), np.uint8)
    Lp[:, :200], Rp[:, :200] = L, R
    with capi.Handle(levels=3) as h:
        a = h.match_full(L, R)
        b = h.match_full(Lp[:, :200], Rp[:, :200])  # np.ascontiguousarray repacks; exercise both paths
    for x, y in zip(a, b):
        assert (x == y).all()


def test_foveated_small(capi, oracle):
    """Foveated mode on a reduced config: 7 levels, fovea_level 3 -> fovea = level-2 dims (319x239)."""
    L, R, _, _ = synth.make_pair(640, 480, 2)
    kw = dict(levels=7, fovea_level=3)
    with capi.Handle(**kw) as h:
        sH, sV, sC, fw, fh = h.match_foveated(L, R)
    wH, wV, wC, ofw, ofh = oracle.match_foveated(L, R, oracle.default_config(**kw))
    assert (fw, fh) == (ofw, ofh) == (319, 239)
    assert sH.shape == (3 * 239, 319)
    _cmp((sH, sV, sC), (wH, wV, wC), "foveated 640x480")


def test_foveated_reference_levels(capi, oracle):
    """Reference defaults (14 levels, foveatelevel 7) on 1232x816: fovea 154x102, stack 714x154."""
    L, R, _, _ = synth.make_pair(1232, 816, 6)
    with capi.Handle(max_iters=4) as h:
        sH, sV, sC, fw, fh = h.match_foveated(L, R)
    wH, wV, wC, ofw, ofh = oracle.match_foveated(L, R, oracle.default_config(max_iters=4))
    assert (fw, fh) == (ofw, ofh)
    assert sH.shape == (7 * fh, fw)
    _cmp((sH, sV, sC), (wH, wV, wC), "foveated 1232x816")


def test_identity_pair_property(capi):
    """L == R: the centre offset wins everywhere in the interior; disparity stays near zero."""
    L, _, _, _ = synth.make_pair(320, 240, 7)
    with capi.Handle(levels=4) as h:
        H, V, C = h.match_full(L, L)
    m = (slice(30, -30), slice(30, -30))
    assert np.abs(H[m]).max() < 0.3 and np.abs(V[m]).max() < 0.3
    assert np.median(np.abs(H[m])) < 0.02 and np.median(np.abs(V[m])) < 0.02
    assert np.median(C[m]) > 0.9


def test_integer_shift_property(capi):
    """right(x) = left(x-3): H converges to +3 (sign: right point = left + disparity)."""
    L, _, _, _ = synth.make_pair(320, 240, 8)
    R = np.roll(L, 3, axis=1)
    with capi.Handle(levels=4) as h:
        H, V, C = h.match_full(L, R)
    m = (slice(30, -30), slice(30, -30))
    assert abs(np.median(H[m]) - 3.0) < 0.25 and abs(np.median(V[m])) < 0.1


def test_determinism_and_handle_reuse(capi):
    """Same handle, different shapes back to back, then the first again: identical bits."""
    A = synth.make_pair(200, 150, 9)[:2]
    B = synth.make_pair(333, 222, 10)[:2]
    with capi.Handle(levels=3) as h:
        r1 = h.match_full(*A)
        h.match_full(*B)
        r2 = h.match_full(*A)
    for x, y in zip(r1, r2):
        assert (x == y).all()


def test_pyramid_and_level_readback(capi, oracle):
    L, R, _, _ = synth.make_pair(320, 240, 11)
    with capi.Handle(levels=4) as h:
        h.match_full(L, R)
        for l in range(4):
            gl, gr = h.pyramid_level(l, 240, 320)
            assert (gl == oracle.pyramid_rgb(L, 4)[l]).all()
            assert (gr == oracle.pyramid_rgb(R, 4)[l]).all()


def test_error_paths(capi):
    with capi.Handle(levels=3) as h:
        L = np.zeros((8, 8, 3), np.uint8)
        with pytest.raises(capi.UgsmError):
            h.match_full(L, np.zeros((8, 9, 3), np.uint8))
        with pytest.raises(capi.UgsmError):  # 2x2 image cannot hold 14 levels
            capi.Handle().match_full(np.zeros((2, 2, 3), np.uint8), np.zeros((2, 2, 3), np.uint8))
        # the handle survives an error
        h.match_full(L, L)


def test_composite_stage_bit_exact(capi, oracle):
    """K4 chain (hierarchicalDisparity) on a random foveated stack vs the oracle."""
    rng = np.random.default_rng(5)
    rows, cols, levels, fl = 480, 640, 7, 3
    w, h = oracle.level_dims(cols, rows, levels)
    fw, fh = w[fl - 1], h[fl - 1]
    sH, sV, sC = (rng.uniform(-3, 3, (fl * fh, fw)).astype(np.float32) for _ in range(3))
    with capi.Handle(levels=levels, fovea_level=fl) as hd:
        got = hd.stage_composite(sH, sV, sC, rows, cols)
    want = oracle.composite(sH, sV, sC, cols, rows, levels, fl)
    _cmp(got, want, "composite stage")


def test_match_composite_fov1(capi, oracle):
    """MatchGPULib::match(L, R, fov = 1): foveated match + hierarchical composite, reduced config and
    reference defaults."""
    L, R, _, _ = synth.make_pair(640, 480, 3)
    kw = dict(levels=7, fovea_level=3)
    with capi.Handle(**kw) as hd:
        got = hd.match_composite(L, R)
    want = oracle.match_composite(L, R, oracle.default_config(**kw))
    _cmp(got, want, "match(fov=1) 640x480")
    L, R, _, _ = synth.make_pair(1232, 816, 6)
    with capi.Handle(max_iters=4) as hd:
        got = hd.match_composite(L, R)
    want = oracle.match_composite(L, R, oracle.default_config(max_iters=4))
    _cmp(got, want, "match(fov=1) 1232x816")


def test_16mp_full_and_foveated_bit_exact(capi, oracle):
    """BASELINE configs[2] and configs[3] at their full size (4928x3264, reference defaults): the oracle needs a
    few seconds on the GPU box's host cores, so the full-size result is compared directly, bit for bit; plus the
    size-independent properties (determinism across handles / graph vs stream launches)."""
    L, R, dH, dV = synth.make_pair(4928, 3264, 0)
    with capi.Handle() as h:
        got = h.match_full(L, R)
        again = h.match_full(L, R)
        fov = h.match_foveated(L, R)
    with capi.Handle(use_graph=0) as h2:
        nograph = h2.match_full(L, R)
    for a, b, c in zip(got, again, nograph):
        assert (a == b).all() and (a == c).all()
    want = oracle.match_full(L, R)
    _cmp(got, want, "16MP full")
    m = (slice(64, -64), slice(64, -64))
    assert np.abs(got[0] - dH)[m].mean() < 0.35 and np.abs(got[1] - dV)[m].mean() < 0.35  # sanity EPE only
    wH, wV, wC, fw, fh = oracle.match_foveated(L, R)
    assert (fw, fh) == (615, 407) == (fov[3], fov[4]) and fov[0].shape == (7 * 407, 615)
    _cmp(fov[:3], (wH, wV, wC), "16MP foveated")


def test_4mp_full_pyramid_bit_exact(capi, oracle):
    """BASELINE configs[1] at its stated size: 2464x1632 (calibrations/calL.xml:5-6), full pyramid, reference
    defaults, against the oracle bit for bit.  Also: the H + V only call (C not copied back) returns the same planes,
    and a pair with black pixels (the levels' range flags clear on the device, K1 runs its guarded march) matches
    the oracle as well."""
    L, R, dH, dV = synth.make_pair(2464, 1632, 1)
    with capi.Handle() as h:
        got = h.match_full(L, R)
        H2 = np.empty_like(got[0])
        V2 = np.empty_like(got[1])
        h.match_full(L, R, out=(H2, V2, None))
    want = oracle.match_full(L, R)
    _cmp(got, want, "4MP full")
    assert np.array_equal(H2, got[0]) and np.array_equal(V2, got[1])
    m = (slice(64, -64), slice(64, -64))
    assert np.abs(got[0] - dH)[m].mean() < 0.35 and np.abs(got[1] - dV)[m].mean() < 0.35  # sanity EPE only
    Lb, Rb = L.copy(), R.copy()
    Lb[200:260, 300:420] = 0
    Rb[700:705, :] = 0
    Rb[1000:1100, 2000:2100] = 0
    with capi.Handle() as h:
        gotb = h.match_full(Lb, Rb)
    _cmp(gotb, oracle.match_full(Lb, Rb), "4MP full with black regions")


def test_pageable_staging_equals_plain_path(capi, monkeypatch):
    """Pageable host buffers of >= 8 MB go through pinned staging with copy threads (what a caller of the
    MatchGPULib facade gets): same bits as the plain cudaMemcpy path, with fresh and with reused output planes,
    also for the foveated stack (too small to be staged on the way out) and for padded rows."""
    L, R, _, _ = synth.make_pair(2048, 1536, 11)
    Lp = np.zeros((1536, 2048 * 3 + 64), np.uint8)
    Rp = np.zeros_like(Lp)
    Lp[:, :2048 * 3] = L.reshape(1536, -1)
    Rp[:, :2048 * 3] = R.reshape(1536, -1)
    Lv = np.lib.stride_tricks.as_strided(Lp, shape=(1536, 2048, 3), strides=(Lp.strides[0], 3, 1))
    Rv = np.lib.stride_tricks.as_strided(Rp, shape=(1536, 2048, 3), strides=(Rp.strides[0], 3, 1))
    monkeypatch.setenv("UGSM_HOST_THREADS", "0")
    with capi.Handle(levels=8) as h:
        want = h.match_full(L, R)
        want_stack = h.match_foveated(L, R)[:3]
    for threads in ("3", "8"):
        monkeypatch.setenv("UGSM_HOST_THREADS", threads)
        with capi.Handle(levels=8) as h:
            got = h.match_full(L, R)
            again = h.match_full(L, R, out=tuple(np.full_like(g, np.nan) for g in got))
            stack = h.match_foveated(L, R)[:3]
            outs = tuple(np.empty((1536, 2048), np.float32) for _ in range(3))
            h._ck(capi.lib().ugsm_match_full(h._h, Lv.ctypes.data_as(capi.C.c_void_p), Rv.ctypes.data_as(capi.C.c_void_p),
                                             1536, 2048, Lp.strides[0], *[o.ctypes.data_as(capi.C.c_void_p) for o in outs]))
        for a, b, c, d in zip(got, again, outs, want):
            assert np.array_equal(a, d) and np.array_equal(b, d) and np.array_equal(c, d)
        for a, d in zip(stack, want_stack):
            assert np.array_equal(a, d)
