"""-m gpu: the streaming strip march k1_stream (ugsm_k1_stream.cuh) against the oracle, bit-exact, through the C-ABI.

The default handle only uses k1_stream above 400 k pixels; these tests force it on small planes (tuning overrides
read when a handle first launches a kernel, i.e. inside the `with` block) so that the oracle finishes in seconds, with short row segments so that
segment seams, the chunked ring and partial strips are all exercised, in both marches: FAST (all image values in
[1, 256]: no division guards) and GENERIC (guarded divisions).
"""
import contextlib
import os

import numpy as np
import pytest

from ug_stereomatcher_b200 import synth

pytestmark = pytest.mark.gpu


@contextlib.contextmanager
def _handle(capi, **env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        h = capi.Handle()
        yield h
        h.close()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _planes(rgb):
    return np.ascontiguousarray(rgb.transpose(2, 0, 1)).astype(np.float32)


def _state(rng, h, w, amp=2.0):
    H = rng.uniform(-amp, amp, (h, w)).astype(np.float32)
    V = rng.uniform(-amp / 2, amp / 2, (h, w)).astype(np.float32)
    C = rng.uniform(0.2, 1.0, (h, w)).astype(np.float32)
    return H, V, C


def _assert_same(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, what
    same = (a == b) | (np.isnan(a) & np.isnan(b))
    if not same.all():
        bad = np.argwhere(~same)
        i = tuple(bad[0])
        raise AssertionError("%s: %d/%d differ, first at %s: gpu=%r oracle=%r" % (what, len(bad), a.size, i, a[i], b[i]))


NAMES = ["Hn", "Vn", "Cn", "srcx", "srcy", "branchH", "branchV"]
STREAM = dict(UGSM_K1_STREAM=1, UGSM_K1_SMALL_PIXELS=0)


@pytest.mark.parametrize("shape", [(4, 5), (16, 30), (17, 31), (33, 121), (100, 250), (118, 159), (61, 241)])
@pytest.mark.parametrize("fast", [1, 0])
@pytest.mark.parametrize("T,first", [(1.0, False), (0.46, False), (1.0, True)])
def test_stream_bit_exact_small_planes(capi, oracle, shape, fast, T, first):
    rows, cols = shape
    rng = np.random.default_rng(rows * 1000 + cols)
    L, R, _, _ = synth.make_pair(cols, rows, 1)
    L3, R3 = _planes(L), _planes(R)
    assert L3.min() >= 1 and R3.min() >= 1  # the FAST precondition holds for the synthetic scenes
    H, V, C = _state(rng, rows, cols)
    want = oracle.correlate(L3, R3, H, V, C, T, first, debug=True)
    # segments of 8 rows (many seams) and of the whole height
    for seg in (8, 4096):
        with _handle(capi, UGSM_K1_FAST=fast, UGSM_K1S_MIN_SEG=seg, UGSM_K1S_SLOTS=100000, **STREAM) as h:
            got = h.stage_correlate(L3, R3, H, V, C, T, first, debug=True)
        for g, w, n in zip(got, want, NAMES):
            _assert_same(g, w, "%s %dx%d T=%g first=%d fast=%d seg=%d" % (n, cols, rows, T, first, fast, seg))


@pytest.mark.parametrize("shape", [(300, 500), (700, 1100)])
def test_stream_large_displacements_and_seams(capi, oracle, shape):
    """Warp offsets far larger than the strip (gathers from anywhere in the right image), default segmentation."""
    rows, cols = shape
    rng = np.random.default_rng(rows + cols)
    L, R, _, _ = synth.make_pair(cols, rows, 4)
    L3, R3 = _planes(L), _planes(R)
    H, V, C = _state(rng, rows, cols, amp=90.0)
    H[5, 7] = np.float32(1e20)    # saturating float -> int conversions of the source index
    V[9, 3] = np.float32(-1e20)
    H[11, 13] = np.float32(np.nan)
    want = oracle.correlate(L3, R3, H, V, C, 0.82, False, debug=True)
    for fast in (1, 0):
        with _handle(capi, UGSM_K1_FAST=fast, **STREAM) as h:
            got = h.stage_correlate(L3, R3, H, V, C, 0.82, False, debug=True)
        for g, w, n in zip(got, want, NAMES):
            _assert_same(g, w, "%s %dx%d fast=%d" % (n, cols, rows, fast))


def test_stream_fast_precondition_extremes(capi, oracle):
    """Image values at the ends of the FAST range [1, 256] (constant planes, checkerboards of the two extremes, random
    values including both ends): the unguarded divisions must still be the IEEE quotients."""
    rows, cols = 96, 150
    rng = np.random.default_rng(11)
    H, V, C = _state(rng, rows, cols)
    yy, xx = np.mgrid[0:rows, 0:cols]
    cases = {
        "ones": (np.ones((3, rows, cols), np.float32), np.ones((3, rows, cols), np.float32)),
        "max": (np.full((3, rows, cols), 256.0, np.float32), np.full((3, rows, cols), 256.0, np.float32)),
        "checker": (np.where((xx + yy) % 2 == 0, 1.0, 256.0).astype(np.float32)[None].repeat(3, 0),
                    np.where((xx // 3 + yy // 2) % 2 == 0, 256.0, 1.0).astype(np.float32)[None].repeat(3, 0)),
        "random": (rng.choice(np.array([1.0, 1.5, 2.0, 17.0, 255.0, 256.0], np.float32), (3, rows, cols)),
                   rng.choice(np.array([1.0, 3.0, 100.0, 256.0], np.float32), (3, rows, cols))),
    }
    for name, (L3, R3) in cases.items():
        L3, R3 = np.ascontiguousarray(L3), np.ascontiguousarray(R3)
        want = oracle.correlate(L3, R3, H, V, C, 1.0, False, debug=True)
        with _handle(capi, UGSM_K1_FAST=1, UGSM_K1S_MIN_SEG=16, UGSM_K1S_SLOTS=100000, **STREAM) as h:
            got = h.stage_correlate(L3, R3, H, V, C, 1.0, False, debug=True)
        for g, w, n in zip(got, want, NAMES):
            _assert_same(g, w, "%s case %s" % (n, name))


@pytest.mark.parametrize("shape", [(64, 90), (300, 500)])
def test_stream_black_regions_take_the_generic_march(capi, oracle, shape):
    """A value below 1 anywhere clears the level's flag on the device: the guarded march runs and matches the oracle
    (0/0 -> NaN -> flat parabola branch, x/0 -> inf -> clamped to 1)."""
    rows, cols = shape
    rng = np.random.default_rng(rows + cols)
    L, R, _, _ = synth.make_pair(cols, rows, 2)
    L3, R3 = _planes(L), _planes(R)
    L3[:, 10:30, 20:60] = 0.0
    R3[:, 25:50, 40:80] = 0.0
    L3[:, rows // 2:rows // 2 + 12, :] = 0.0
    R3[:, rows // 2:rows // 2 + 12, :] = 0.0
    L3[1, 40:44, 5:9] = 1e-12
    R3[1, 40:44, 5:9] = 1e-12
    H, V, C = _state(rng, rows, cols)
    want = oracle.correlate(L3, R3, H, V, C, 1.0, False, debug=True)
    with _handle(capi, UGSM_K1_FAST=1, UGSM_K1S_MIN_SEG=16, UGSM_K1S_SLOTS=100000, **STREAM) as h:
        got = h.stage_correlate(L3, R3, H, V, C, 1.0, False, debug=True)
    for g, w, n in zip(got, want, NAMES):
        _assert_same(g, w, "%s %dx%d with black regions" % (n, cols, rows))


def test_stream_equals_tiled_kernel_at_4mp(capi):
    """2464x1632 (BASELINE configs[1] shape): k1_stream and the tiled k1_correlate produce identical planes."""
    rows, cols = 1632, 2464
    rng = np.random.default_rng(3)
    L, R, _, _ = synth.make_pair(cols, rows, 5)
    L3, R3 = _planes(L), _planes(R)
    H, V, C = _state(rng, rows, cols, amp=6.0)
    outs = []
    for env in (dict(UGSM_K1_STREAM=1, UGSM_K1_FAST=1), dict(UGSM_K1_STREAM=1, UGSM_K1_FAST=0), dict(UGSM_K1_STREAM=0)):
        with _handle(capi, **env) as h:
            outs.append(h.stage_correlate(L3, R3, H, V, C, 0.64, False))
    for k, n in enumerate(["Hn", "Vn", "Cn"]):
        _assert_same(outs[0][k], outs[2][k], "%s FAST stream vs tiled" % n)
        _assert_same(outs[1][k], outs[2][k], "%s GENERIC stream vs tiled" % n)
