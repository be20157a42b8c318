"""CPU: the point-cloud oracle (oracle/ugsm_oracle_cloud.c) against the reference's own code.

Pins: (1) tests/golden/refpc_cloud.npz = outputs of the UNMODIFIED reference source
src/pointcloud/getPointCloud.cpp compiled through stubs (oracle/_ref/refpc); (2) when that binary is
present (it is built wherever /root/reference is mounted and travels with the repo snapshot), a live
comparison on fresh random and adversarial inputs; (3) hand-derived known answers.
All comparisons are on the raw float bits.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_refpc_golden import cameras, cloud_cases, fovea_dims  # noqa: E402

REFPC = os.path.join(ROOT, "oracle", "_ref", "refpc")


def oracle_cloud(o, c, stride=4):
    if c["mode"] == 0:
        return o.cloud_full(c["P1"], c["P2"], c["dX"], c["dY"], c["img"], bgr=True, sampling=c["sampling"],
                            stride=stride)
    return o.cloud_foveated(c["P1"], c["P2"], c["dX"], c["dY"], c["fov_h"], c["src_level"], c["img"].shape[0],
                            c["img"].shape[1], c["img"], bgr=True, sampling=c["sampling"], stride=stride)


def test_golden_matches_reference_outputs(oracle):
    gold = np.load(os.path.join(ROOT, "tests", "golden", "refpc_cloud.npz"))
    cases = cloud_cases()
    assert sorted(gold.files) == sorted(c["name"] for c in cases)
    for c in cases:
        got = oracle_cloud(oracle, c).view(np.uint32)
        assert got.shape == gold[c["name"]].shape, c["name"]
        assert np.array_equal(got, gold[c["name"]]), c["name"]


@pytest.mark.skipif(not os.path.exists(REFPC), reason="oracle/_ref/refpc not built (needs /root/reference)")
def test_live_against_reference_binary(oracle, tmp_path):
    rng = np.random.default_rng(77)
    total = 0
    for t in range(6):
        rows, cols, s = int(rng.integers(5, 80)), int(rng.integers(5, 110)), int(rng.integers(1, 4))
        P1, P2 = cameras(rng, cols, rows)
        # large, negative and zero disparities, including exact integers
        dX = (rng.normal(0, 40, (rows, cols)) * (rng.random((rows, cols)) < 0.8)).astype(np.float32)
        dY = np.round(rng.normal(0, 3, (rows, cols))).astype(np.float32)
        img = rng.integers(0, 256, (rows, cols, 3), dtype=np.uint8)
        ref = oracle.run_refpc(0, P1, P2, dX, dY, img, sampling=s, workdir=str(tmp_path))
        mine = oracle.cloud_full(P1, P2, dX, dY, img, bgr=True, sampling=s)
        assert np.array_equal(ref.view(np.uint32), mine.view(np.uint32))
        total += ref.shape[0]
    for t in range(4):
        im_rows, im_cols = int(rng.integers(300, 700)), int(rng.integers(400, 900))
        fh, fw = fovea_dims(im_rows, im_cols)
        P1, P2 = cameras(rng, im_cols, im_rows)
        sX = rng.normal(0, 6, (7 * fh, fw)).astype(np.float32)
        sY = rng.normal(0, 2, (7 * fh, fw)).astype(np.float32)
        img = rng.integers(0, 256, (im_rows, im_cols, 3), dtype=np.uint8)
        lev, s = int(rng.integers(0, 7)), int(rng.integers(1, 3))
        ref = oracle.run_refpc(1, P1, P2, sX, sY, img, sampling=s, src_level=lev, workdir=str(tmp_path))
        mine = oracle.cloud_foveated(P1, P2, sX, sY, fh, lev, im_rows, im_cols, img, sampling=s)
        assert np.array_equal(ref.view(np.uint32), mine.view(np.uint32))
        total += ref.shape[0]
    assert total > 10000


def oracle_range(o, c):
    if c["mode"] == 0:
        return o.range_full(c["P1"], c["P2"], c["dX"], c["dY"])
    return o.range_foveated(c["P1"], c["P2"], c["dX"], c["dY"], c["fov_h"], c["src_level"], c["img"].shape[0],
                            c["img"].shape[1])


def test_range_golden_matches_reference_outputs(oracle):
    """Range maps (getRangePoint / getRangePointFOV, getPointCloud.cpp:951-1019): committed outputs of the unmodified
    reference source (tests/golden/refpc_range.npz, refpc modes 2 / 3)."""
    gold = np.load(os.path.join(ROOT, "tests", "golden", "refpc_range.npz"))
    for c in cloud_cases():
        got = oracle_range(oracle, c)
        assert np.array_equal(got.view(np.uint32).ravel(), gold[c["name"]]), c["name"]
        # and it is the z column of the cloud of the same inputs (the two functions share their Z expression)
        if c["sampling"] == 1:
            cloud = oracle_cloud(oracle, c)
            z = cloud[:, 2].reshape(got.shape[1], got.shape[0]).T  # cloud order is column-major
            assert np.array_equal(z.view(np.uint32), got.view(np.uint32)), c["name"]


@pytest.mark.skipif(not os.path.exists(REFPC), reason="oracle/_ref/refpc not built (needs /root/reference)")
def test_range_live_against_reference_binary(oracle, tmp_path):
    rng = np.random.default_rng(78)
    for t in range(4):
        rows, cols = int(rng.integers(5, 80)), int(rng.integers(5, 110))
        P1, P2 = cameras(rng, cols, rows)
        dX = (rng.normal(0, 40, (rows, cols)) * (rng.random((rows, cols)) < 0.8)).astype(np.float32)
        dY = np.round(rng.normal(0, 3, (rows, cols))).astype(np.float32)
        img = np.zeros((rows, cols, 3), np.uint8)
        ref = oracle.run_refpc(2, P1, P2, dX, dY, img, workdir=str(tmp_path)).reshape(rows, cols)
        assert np.array_equal(ref.view(np.uint32), oracle.range_full(P1, P2, dX, dY).view(np.uint32))
    for t in range(3):
        im_rows, im_cols = int(rng.integers(300, 700)), int(rng.integers(400, 900))
        fh, fw = fovea_dims(im_rows, im_cols)
        P1, P2 = cameras(rng, im_cols, im_rows)
        sX = rng.normal(0, 6, (7 * fh, fw)).astype(np.float32)
        sY = rng.normal(0, 2, (7 * fh, fw)).astype(np.float32)
        img = np.zeros((im_rows, im_cols, 3), np.uint8)
        lev = int(rng.integers(0, 7))
        ref = oracle.run_refpc(3, P1, P2, sX, sY, img, src_level=lev, workdir=str(tmp_path)).reshape(fh, fw)
        mine = oracle.range_foveated(P1, P2, sX, sY, fh, lev, im_rows, im_cols)
        assert np.array_equal(ref.view(np.uint32), mine.view(np.uint32))


def test_rectified_pair_known_answer(oracle):
    """P1 = K[I|0], P2 = K[I|(-B,0,0)]: disparity d = x2 - x1 = -f*B/Z, so the closed form must return
    Z = -f*B/d, X = (x - cx) Z / f, Y = (y - cy) Z / f (to float accuracy; the formula is exact algebra)."""
    f, cx, cy, B = 1000.0, 320.0, 240.0, 0.12
    P1 = [f, 0, cx, 0, 0, f, cy, 0, 0, 0, 1, 0]
    P2 = [f, 0, cx, -f * B, 0, f, cy, 0, 0, 0, 1, 0]
    for (x, y, Z) in [(100, 50, 2.0), (320, 240, 0.8), (600, 400, 5.0)]:
        d = -f * B / Z
        p = oracle.triangulate(P1, P2, x, y, x + d, y)
        want = np.array([(x - cx) * Z / f, (y - cy) * Z / f, Z])
        assert np.allclose(p, want, rtol=2e-4, atol=1e-4), (p, want)


def test_margins_and_scale_known_answers(oracle):
    # 16 MP: the fovea of stack level 0 is level 6 of the pyramid (615x407) centred in 4928x3264
    assert oracle.cloud_margins(3264, 4928, 7, 0) == (4928 // 2 - 615 // 2, 3264 // 2 - 407 // 2) == (2157, 1429)
    # stack level 6 covers the whole image
    assert oracle.cloud_margins(3264, 4928, 7, 6) == (0, 0)
    assert oracle.cloud_margins(3264, 4928, 7, 1) == (4928 // 2 - 870 // 2, 3264 // 2 - 576 // 2)
    assert oracle.cloud_scale(0) == 1.0
    assert oracle.cloud_scale(2) == pytest.approx(2.0, rel=1e-6)
    assert np.float32(oracle.cloud_scale(1)) == np.float32(1.41421356237309504880)


def test_point_order_stride_and_colour(oracle):
    rng = np.random.default_rng(5)
    rows, cols, s = 7, 9, 2
    P1, P2 = cameras(rng, cols, rows)
    dX = rng.normal(0, 2, (rows, cols)).astype(np.float32)
    dY = rng.normal(0, 1, (rows, cols)).astype(np.float32)
    img = rng.integers(0, 256, (rows, cols, 3), dtype=np.uint8)
    p4 = oracle.cloud_full(P1, P2, dX, dY, img, bgr=True, sampling=s, stride=4)
    p8 = oracle.cloud_full(P1, P2, dX, dY, img, bgr=True, sampling=s, stride=8)
    ncol, nrow = (cols + s - 1) // s, (rows + s - 1) // s
    assert p4.shape == (ncol * nrow, 4) and oracle.cloud_count(rows, cols, s) == ncol * nrow
    # column-major: record ci * nrow + rj is pixel (x, y) = (ci * s, rj * s)
    for ci in range(ncol):
        for rj in range(nrow):
            rec = p4[ci * nrow + rj]
            x, y = ci * s, rj * s
            want = oracle.triangulate(P1, P2, x, y, np.float32(x) + dX[y, x], np.float32(y) + dY[y, x])
            assert np.array_equal(rec[:3].view(np.uint32), want.view(np.uint32))
            b, g, r = (int(v) for v in img[y, x])
            assert int(rec[3:4].view(np.uint32)[0]) == (r << 16 | g << 8 | b)
    # pcl::PointXYZRGB layout
    assert np.array_equal(p8[:, :3].view(np.uint32), p4[:, :3].view(np.uint32))
    assert np.all(p8[:, 3] == 1.0) and np.array_equal(p8[:, 4].view(np.uint32), p4[:, 3].view(np.uint32))
    assert np.all(p8[:, 5:].view(np.uint32) == 0)
    # an RGB-ordered image gives the same colours with bgr=False
    p_rgb = oracle.cloud_full(P1, P2, dX, dY, img[:, :, ::-1], bgr=False, sampling=s)
    assert np.array_equal(p_rgb.view(np.uint32), p4.view(np.uint32))


@pytest.mark.skipif(not os.path.exists(REFPC), reason="oracle/_ref/refpc not built (needs /root/reference)")
def test_reference_binary_timing(oracle, tmp_path):
    """The reference's own per-pixel host loop (single thread, like the node) on a 1 MP plane, scaled to 16 MP:
    the CPU figure quoted beside the K5 kernel time in profiles/README.md (run with -s to see it)."""
    import time

    rows, cols = 816, 1232
    rng = np.random.default_rng(3)
    P1, P2 = cameras(rng, cols, rows)
    dX = rng.normal(0, 8, (rows, cols)).astype(np.float32)
    dY = rng.normal(0, 1, (rows, cols)).astype(np.float32)
    img = rng.integers(0, 256, (rows, cols, 3), dtype=np.uint8)
    t0 = time.perf_counter()
    ref = oracle.run_refpc(0, P1, P2, dX, dY, img, workdir=str(tmp_path))
    dt = time.perf_counter() - t0
    assert ref.shape == (rows * cols, 4)
    print("reference point-cloud loop: %.2f s per 1 MP incl. file I/O -> %.1f s per 16 MP cloud" % (dt, dt * 16))
