"""-m gpu: the row-pipelined smoothing kernel k2_stream (ugsm_k2_stream.cuh) against the oracle, bit-exact, through
the C-ABI.  The default handle only uses it above 700 k pixels; these tests force it on small planes with short row
segments (tuning overrides read when a handle first launches a kernel) so that strip seams, segment seams, the
pipeline fill / drain and the image borders are all exercised while the oracle finishes in seconds."""
import contextlib
import os

import numpy as np
import pytest

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


def _state(rng, h, w, amp=2.0):
    H = rng.uniform(-amp, amp, (h, w)).astype(np.float32)
    V = rng.uniform(-amp / 2, amp / 2, (h, w)).astype(np.float32)
    C = rng.uniform(0.2, 1.0, (h, w)).astype(np.float32)
    return H, V, C


def _assert_same(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    same = (a == b) | (np.isnan(a) & np.isnan(b))
    if not same.all():
        bad = np.argwhere(~same)
        i = tuple(bad[0])
        raise AssertionError("%s: %d/%d differ, first at %s: gpu=%r oracle=%r" % (what, len(bad), a.size, i, a[i], b[i]))


STREAM = dict(UGSM_K2_STREAM=1, UGSM_K2_SMALL_PIXELS=0, UGSM_K2S_MIN_PIXELS=0, UGSM_K2S_MAX_PASSES=10)


@pytest.mark.parametrize("shape", [(4, 5), (32, 64), (33, 65), (100, 250), (118, 159), (61, 241), (300, 500)])
@pytest.mark.parametrize("passes", [0, 1, 5, 6, 10])
def test_k2_stream_bit_exact(capi, oracle, shape, passes):
    rows, cols = shape
    rng = np.random.default_rng(7 + rows + cols + passes)
    H, V, C = _state(rng, rows, cols)
    want = oracle.smooth(H, V, C, passes)
    for seg in (8, 4096):
        with _handle(capi, UGSM_K2S_MIN_SEG=seg, UGSM_K2S_SLOTS=100000, **STREAM) as h:
            got = h.stage_smooth(H, V, C, passes)
        for g, w, n in zip(got, want, "HVC"):
            _assert_same(g, w, "smooth %s %dx%d S=%d seg=%d" % (n, cols, rows, passes, seg))


def _assert_same_away_from_nonfinite(g, w, what):
    """Bit equality everywhere except within 2 px of a non-finite value: the oracle multiplies the box filter's zero
    outer taps (0 * NaN = NaN, MatchLib.cu:1616-1623), the CUDA kernels skip them (same rule as
    tests/test_gpu_stages.py::test_smooth_special_values_bit_exact)."""
    bad = ~np.isfinite(w) | ~np.isfinite(g)
    near = bad.copy()
    for d in (1, 2):
        near[d:, :] |= bad[:-d, :]
        near[:-d, :] |= bad[d:, :]
        near[:, d:] |= bad[:, :-d]
        near[:, :-d] |= bad[:, d:]
    ok = (g == w) | near
    assert ok.all(), "%s: %d px differ away from non-finite values" % (what, int((~ok).sum()))
    assert (~near).mean() > 0.5, what


def test_k2_stream_zero_and_special_values(capi, oracle):
    """Zero disparities / confidences (0/0), huge and tiny values take the guarded division's slow path."""
    rows, cols = 90, 300
    rng = np.random.default_rng(3)
    H, V, C = _state(rng, rows, cols)
    H[10:30, 20:200] = 0.0
    V[10:30, 20:200] = 0.0
    C[40:60, 100:260] = 0.0
    H[70:75, 5:50] = np.float32(1e30)
    V[70:75, 5:50] = np.float32(-1e-30)
    C[80:85, 200:290] = np.float32(1e-25)
    want = oracle.smooth(H, V, C, 5)
    with _handle(capi, UGSM_K2S_MIN_SEG=16, UGSM_K2S_SLOTS=100000, **STREAM) as h:
        got = h.stage_smooth(H, V, C, 5)
    for g, w, n in zip(got, want, "HVC"):
        _assert_same_away_from_nonfinite(g, w, "smooth specials %s" % n)


def test_k2_stream_equals_tiled_kernel_at_4mp(capi):
    """2464x1632 (BASELINE configs[1] shape), S = 10 and 5: k2_stream and the tiled k2_smooth_box agree bit for bit."""
    rows, cols = 1632, 2464
    rng = np.random.default_rng(5)
    H, V, C = _state(rng, rows, cols, amp=8.0)
    for passes in (10, 5):
        outs = []
        for env in (dict(UGSM_K2_STREAM=1, UGSM_K2S_MAX_PASSES=10), dict(UGSM_K2_STREAM=0)):
            with _handle(capi, **env) as h:
                outs.append(h.stage_smooth(H, V, C, passes))
        for k, n in enumerate("HVC"):
            _assert_same(outs[0][k], outs[1][k], "%s S=%d stream vs tiled" % (n, passes))
