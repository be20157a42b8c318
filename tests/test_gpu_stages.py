"""-m gpu: every fused kernel against the oracle on the same seeded inputs, bit-exact.

All calls go through the C-ABI (ug_stereomatcher_b200/libugsm.so via ctypes).
"""
import numpy as np
import pytest

from ug_stereomatcher_b200 import synth

pytestmark = pytest.mark.gpu


def _state(rng, h, w, amp=2.0):
    H = rng.uniform(-amp, amp, (h, w)).astype(np.float32)
    V = rng.uniform(-amp / 2, amp / 2, (h, w)).astype(np.float32)
    C = rng.uniform(0.2, 1.0, (h, w)).astype(np.float32)
    return H, V, C


def _planes(rgb):
    return np.ascontiguousarray(rgb.transpose(2, 0, 1)).astype(np.float32)


def _assert_same(a, b, what):
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, what
    same = (a == b) | (np.isnan(a) & np.isnan(b))
    if not same.all():
        bad = np.argwhere(~same)
        i = tuple(bad[0])
        raise AssertionError("%s: %d/%d differ, first at %s: gpu=%r oracle=%r, max|d|=%g" % (
            what, len(bad), a.size, i, a[i], b[i], np.nanmax(np.abs(a.astype(np.float64) - b))))


@pytest.mark.parametrize("shape", [(480, 640), (407, 615), (34, 53), (4, 5), (61, 241), (16, 30), (17, 31)])
def test_pyramid_bit_exact(gpu, oracle, capi, shape):
    rows, cols = shape
    levels = 14 if min(rows, cols) >= 400 else (6 if min(rows, cols) >= 30 else 2)
    L, _, _, _ = synth.make_pair(cols, rows, 3)
    with capi.Handle(levels=levels) as h:
        got = h.stage_pyramid(L)
    want = oracle.pyramid_rgb(L, levels)
    for l in range(levels):
        _assert_same(got[l], want[l], "pyramid level %d of %dx%d" % (l, cols, rows))


@pytest.mark.parametrize("shape", [(911, 1203), (1021, 1534), (900, 2050)])
def test_pyramid_large_planes_bit_exact(gpu, oracle, capi, shape):
    """Planes above the size where K0 switches to the register-marching kernels (vector loads, running column
    sums), at widths that are not multiples of 4 / 64 / 128 so that border lanes and partial warps are hit."""
    rows, cols = shape
    L, _, _, _ = synth.make_pair(cols, rows, 7)
    with capi.Handle(levels=6) as h:
        got = h.stage_pyramid(L)
    want = oracle.pyramid_rgb(L, 6)
    for l in range(6):
        _assert_same(got[l], want[l], "pyramid level %d of %dx%d" % (l, cols, rows))


def test_correlate_large_plane_bit_exact(gpu, oracle):
    """K1 on a 1.1 MP plane: the left-energy planes U come from the dense register-marching K0 kernel."""
    rows, cols = 911, 1203
    rng = np.random.default_rng(5)
    L, R, _, _ = synth.make_pair(cols, rows, 2)
    L3, R3 = _planes(L), _planes(R)
    H, V, C = _state(rng, rows, cols)
    got = gpu.stage_correlate(L3, R3, H, V, C, 0.64, False)
    want = oracle.correlate(L3, R3, H, V, C, 0.64, False)
    for g, w, n in zip(got, want, ["Hn", "Vn", "Cn"]):
        _assert_same(g, w, "correlate %s on %dx%d" % (n, cols, rows))


@pytest.mark.parametrize("shape", [(118, 159), (34, 53), (4, 5), (16, 30), (17, 31), (33, 121), (100, 250)])
@pytest.mark.parametrize("T,first", [(1.0, False), (0.46, False), (1.0, True)])
def test_correlate_bit_exact(gpu, oracle, shape, T, first):
    rows, cols = shape
    rng = np.random.default_rng(rows * 1000 + cols)
    L, R, _, _ = synth.make_pair(cols, rows, 1)
    L3, R3 = _planes(L), _planes(R)
    H, V, C = _state(rng, rows, cols)
    got = gpu.stage_correlate(L3, R3, H, V, C, T, first, debug=True)
    want = oracle.correlate(L3, R3, H, V, C, T, first, debug=True)
    names = ["Hn", "Vn", "Cn", "srcx", "srcy", "branchH", "branchV"]
    for g, w, n in zip(got, want, names):
        _assert_same(g, w, "%s %dx%d T=%g first=%d" % (n, cols, rows, T, first))


@pytest.mark.parametrize("shape", [(118, 159), (34, 53), (4, 5), (32, 64), (33, 65), (100, 250)])
@pytest.mark.parametrize("passes", [0, 1, 5, 10])
def test_smooth_bit_exact(gpu, oracle, shape, passes):
    rows, cols = shape
    rng = np.random.default_rng(7 + rows + cols + passes)
    H, V, C = _state(rng, rows, cols)
    got = gpu.stage_smooth(H, V, C, passes)
    want = oracle.smooth(H, V, C, passes)
    for g, w, n in zip(got, want, "HVC"):
        _assert_same(g, w, "smooth %s %dx%d S=%d" % (n, cols, rows, passes))


@pytest.mark.parametrize("src,fine", [((34, 53), (49, 75)), ((407, 615), (576, 870)), ((4, 5), (6, 8))])
def test_upsample_bit_exact(gpu, oracle, src, fine):
    rng = np.random.default_rng(11)
    H, V, C = _state(rng, *src)
    got = gpu.stage_upsample(H, V, C, fine[1], fine[0])
    want = oracle.upsample_state(H, V, C, fine[1], fine[0])
    for g, w, n in zip(got, want, "HVC"):
        _assert_same(g, w, "upsample %s" % n)


def test_upsample_crop_bit_exact(gpu, oracle):
    """Foveated variant: 615x407 -> 870x576 then centre crop 615x407 at (128, 85) (MatchGPULib.cpp:1612-1615)."""
    rng = np.random.default_rng(12)
    H, V, C = _state(rng, 407, 615)
    args = (870, 576, 870 // 2 - 615 // 2, 576 // 2 - 407 // 2, 615, 407)
    assert args[2:4] == (128, 85)
    got = gpu.stage_upsample(H, V, C, *args)
    want = oracle.upsample_state(H, V, C, *args)
    for g, w, n in zip(got, want, "HVC"):
        _assert_same(g, w, "upsample+crop %s" % n)


@pytest.mark.parametrize("shape,iters,passes,coarsest", [((58, 79), 22, 5, True), ((118, 159), 10, 5, False),
                                                         ((83, 112), 4, 10, False)])
def test_match_level_bit_exact(gpu, oracle, shape, iters, passes, coarsest):
    rows, cols = shape
    rng = np.random.default_rng(5)
    L, R, _, _ = synth.make_pair(cols, rows, 2)
    L3, R3 = _planes(L), _planes(R)
    if coarsest:
        H = np.zeros((rows, cols), np.float32)
        V, C = H.copy(), H.copy()
    else:
        H, V, C = _state(rng, rows, cols, amp=1.0)
    got = gpu.stage_match_level(L3, R3, H, V, C, iters, passes, coarsest)
    want = oracle.match_level(L3, R3, H, V, C, iters, passes, coarsest)
    for g, w, n in zip(got, want, "HVC"):
        _assert_same(g, w, "match_level %s %dx%d" % (n, cols, rows))


def test_match_level_bit_exact_vs_live_reference_cuda(gpu, tmp_path):
    """This repo's CUDA path vs the UNMODIFIED reference CUDA code (oracle/_ref/refcuda_stage, built from
    /root/reference by oracle/refbuild/) run live on the same GPU, at a tile-aligned size where the
    reference has no out-of-bounds accidents: bit-exact, no oracle in between."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "oracle", "_ref", "refcuda_stage")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/refcuda_stage not built (reference checkout absent at build time)")
    sys.path.insert(0, os.path.join(root, "tests", "golden"))
    import make_refcuda_stage_golden as gen

    L3, R3, H0, V0, C0 = gen.stage_inputs()
    W, H = gen.W, gen.H
    fin, fout = str(tmp_path / "in.f32"), str(tmp_path / "out.f32")
    for name, level, iters, passes, coarsest in gen.LEVEL_CASES:
        if coarsest:
            h0 = np.zeros_like(H0)
            v0, c0 = h0.copy(), h0.copy()
        else:
            h0, v0, c0 = H0, V0, C0
        np.concatenate([L3.ravel(), R3.ravel(), h0.ravel(), v0.ravel(), c0.ravel()]).astype(np.float32).tofile(fin)
        p = subprocess.run([exe, "level", str(W), str(H), str(level), fin, fout], stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, timeout=600)
        assert p.returncode == 0, p.stdout.decode(errors="ignore")[-2000:]
        ref = np.fromfile(fout, np.float32).reshape(3, H, W)
        got = gpu.stage_match_level(L3, R3, h0, v0, c0, iters, passes, coarsest)
        for g, r, n in zip(got, ref, "HVC"):
            _assert_same(g, r, "%s %s vs live reference CUDA" % (name, n))


@pytest.mark.parametrize("level,iters,passes", [(13, 2, 10), (11, 6, 5)])
def test_large_kernels_bit_exact_vs_live_reference_cuda(capi, tmp_path, level, iters, passes):
    """The LARGE-level kernels (k1_stream with its TMA tile loads; the row-pipelined k2_stream at S = 10 and at
    S = 5) against the UNMODIFIED reference CUDA code run live on the same GPU at 1024x768 (W % 128 = 0,
    H % 64 = 0: no out-of-bounds accidents in the reference, 786 k pixels: above both switch points): the finest-level
    schedule (reference level 13: 2 iterations, S = 10) and the third-finest (level 11: 6 iterations, S = 5, threshold
    schedule), bit-exact with no oracle in between, in the FAST and in the GENERIC march of K1."""
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "oracle", "_ref", "refcuda_stage")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/refcuda_stage not built (reference checkout absent at build time)")
    W, H = 1024, 768
    L, R, dH, dV = synth.make_pair(W, H, 6)
    L3, R3 = _planes(L), _planes(R)
    rng = np.random.default_rng(31)
    H0 = (0.8 * dH + rng.uniform(-0.3, 0.3, dH.shape)).astype(np.float32)
    V0 = (0.8 * dV + rng.uniform(-0.3, 0.3, dV.shape)).astype(np.float32)
    C0 = rng.uniform(0.4, 1.0, dH.shape).astype(np.float32)
    fin, fout = str(tmp_path / "in.f32"), str(tmp_path / "out.f32")
    np.concatenate([L3.ravel(), R3.ravel(), H0.ravel(), V0.ravel(), C0.ravel()]).astype(np.float32).tofile(fin)
    p = subprocess.run([exe, "level", str(W), str(H), str(level), fin, fout], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, timeout=900)
    assert p.returncode == 0, p.stdout.decode(errors="ignore")[-2000:]
    ref = np.fromfile(fout, np.float32).reshape(3, H, W)
    old = {k: os.environ.get(k) for k in ("UGSM_K1_FAST", "UGSM_K2S_MIN_PIXELS")}
    try:
        for fast in ("1", "0"):
            os.environ["UGSM_K1_FAST"] = fast
            os.environ["UGSM_K2S_MIN_PIXELS"] = "0"  # S = 5 at 786 k pixels: run the pipelined K2 as well
            with capi.Handle() as h:
                got = h.stage_match_level(L3, R3, H0, V0, C0, iters, passes, False)
            for g, r, n in zip(got, ref, "HVC"):
                _assert_same(g, r, "1024x768 level %d %s (K1 fast=%s) vs live reference CUDA" % (level, n, fast))
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _assert_same_away_from_nonfinite(g, w, what):
    """Bit equality everywhere except within 2 px of a non-finite value: the oracle multiplies the box
    filter's zero outer taps (0 * NaN = NaN, MatchLib.cu:1616-1623), the CUDA kernel skips them.  State
    planes are never non-finite in the matcher; this only isolates the division."""
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


def test_smooth_special_values_bit_exact(gpu, oracle):
    """The smoothing divisions use a guarded fast path with a shared reciprocal; zeros, denormals, values
    outside [2^-62, 2^62] and zero denominators must take the full IEEE path and still match the oracle."""
    rng = np.random.default_rng(99)
    rows, cols = 70, 150
    H, V, C = _state(rng, rows, cols)
    specials = np.array([0.0, -0.0, 1e-45, -1e-42, 1e-38, 3e-30, -7e-25, 1e-20, 1e19, -3e25, 2.0 ** -62, 2.0 ** 62,
                         2.0 ** -63, 2.0 ** 63], np.float32)
    for plane in (H, V):
        ys = rng.integers(0, rows, 400)
        xs = rng.integers(0, cols, 400)
        plane[ys, xs] = specials[rng.integers(0, len(specials), 400)]
    H[10:20, 30:60] = 0.0          # zero numerators (the coarsest level starts like this)
    C[40:44, 100:120] = 0.0        # zero denominators -> 0/0 and x/0
    C[50:52, 10:30] = 1e-30        # tiny denominators
    C[5:7, 5:40] = 3e20            # huge denominators
    for passes in (1, 5):
        got = gpu.stage_smooth(H, V, C, passes)
        want = oracle.smooth(H, V, C, passes)
        for g, w, n in zip(got, want, "HVC"):
            _assert_same_away_from_nonfinite(g, w, "smooth specials %s S=%d" % (n, passes))
    # same through the small-level kernel (16x16 tiles, one thread per cell) and the big one
    for shape in ((40, 60), (700, 1100)):
        H2, V2, C2 = _state(rng, *shape)
        ys = rng.integers(0, shape[0], 300)
        xs = rng.integers(0, shape[1], 300)
        H2[ys, xs] = specials[rng.integers(0, len(specials), 300)]
        C2[ys[:50], xs[:50]] = 0.0
        got = gpu.stage_smooth(H2, V2, C2, 5)
        want = oracle.smooth(H2, V2, C2, 5)
        for g, w, n in zip(got, want, "HVC"):
            _assert_same_away_from_nonfinite(g, w, "smooth specials %s %s" % (n, shape))


@pytest.mark.parametrize("shape", [(64, 90), (300, 500), (700, 1100)])
def test_correlate_black_and_dark_regions_bit_exact(gpu, oracle, shape):
    """The NCC^2 divisions use a grouped, guarded fast path.  Black regions (0/0 -> NaN -> flat parabola branch,
    x/0 -> inf -> clamped to 1), very dark regions and a saturated region must take the right path and match the
    oracle bit for bit in both K1 variants (small / large levels)."""
    rows, cols = shape
    rng = np.random.default_rng(rows + cols)
    L, R, _, _ = synth.make_pair(cols, rows, 2)
    L3, R3 = _planes(L), _planes(R)
    L3[:, 10:30, 20:60] = 0.0           # black in the left image only
    R3[:, 25:50, 40:80] = 0.0           # black in the right image, partly overlapping
    L3[:, rows // 2:rows // 2 + 12, :] = 0.0
    R3[:, rows // 2:rows // 2 + 12, :] = 0.0   # black in both
    L3[:, -20:, -40:] = 1.0             # very dark
    R3[:, -20:, -40:] = 1.0
    L3[:, :8, :8] = 255.0
    R3[:, :8, :8] = 255.0
    L3[1, 40:44, 5:9] = 1e-12           # tiny (a blurred, almost black coarse level)
    R3[1, 40:44, 5:9] = 1e-12
    H, V, C = _state(rng, rows, cols)
    got = gpu.stage_correlate(L3, R3, H, V, C, 1.0, False, debug=True)
    want = oracle.correlate(L3, R3, H, V, C, 1.0, False, debug=True)
    for g, w, n in zip(got, want, ["Hn", "Vn", "Cn", "srcx", "srcy", "branchH", "branchV"]):
        _assert_same(g, w, "%s %dx%d with black regions" % (n, cols, rows))


@pytest.mark.parametrize("T", [1.0, 0.55, 0.1])
def test_poly_fit_exhaustive_bit_exact(gpu, oracle, T):
    """K1's parabola fit replaces the reference's double division by a guarded f32 fast path (double rounding is
    innocuous for f32 operands) -- compare millions of triples, including ties, tiny/huge/zero/NaN values and the
    rescale branch, with the oracle's literal double arithmetic."""
    rng = np.random.default_rng(int(T * 100))
    n = 2_000_000
    l, c, r = (rng.uniform(0, 1, n).astype(np.float32) for _ in range(3))
    # near-flat and near-symmetric cases, where c1 ~ 0 and b1 ~ 0
    k = n // 4
    l[:k] = c[:k] + rng.normal(0, 1e-6, k).astype(np.float32)
    r[:k] = c[:k] + rng.normal(0, 1e-6, k).astype(np.float32)
    r[k:2 * k] = l[k:2 * k]
    # peaks above 1 (rescale branch): centre near 1, neighbours close
    c[2 * k:2 * k + 50000] = 1.0
    l[2 * k:2 * k + 50000] = rng.uniform(0.9, 1.0, 50000)
    r[2 * k:2 * k + 50000] = rng.uniform(0.9, 1.0, 50000)
    specials = np.array([0.0, 1.0, 1e-45, 1e-40, 1e-38, 1e-30, 1e-20, 1e-10, 0.5, np.nan, 2.0 ** -62, 2.0 ** -63,
                         2.0 ** -126, 3e-39], np.float32)
    idx = rng.integers(0, n, 300000)
    for arr in (l, c, r):
        arr[idx[rng.integers(0, len(idx), 100000)]] = specials[rng.integers(0, len(specials), 100000)]
    got = gpu.stage_poly(l, c, r, T)
    want = oracle.poly_n(l, c, r, T)
    for g, w, name in zip(got, want, ["offset", "conf", "branch"]):
        _assert_same(g, w, "poly %s T=%g" % (name, T))
    assert (want[2] == 2).sum() > 1000 and (want[2] == 1).sum() > 100000 and (want[2] == 0).sum() > 100000
