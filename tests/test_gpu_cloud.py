"""-m gpu: the point-cloud kernel (K5) through the C-ABI against the oracle, bit for bit.

The oracle itself is pinned against the unmodified reference source (tests/test_cloud_oracle.py).
Integer outputs (packed colour, point order / count) and the float coordinates are compared on
their raw bits.  Inputs are finite: for non-finite disparities the reference's float -> int
conversion (getPointCloud.cpp:900-901) is undefined behaviour and is not part of the contract.
"""
import os
import sys

import numpy as np
import pytest

from ug_stereomatcher_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_refpc_golden import cameras, cloud_cases, fovea_dims  # noqa: E402

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def same(got, want, what):
    assert got.shape == want.shape, what
    ne = bits(got) != bits(want)
    assert not ne.any(), "%s: %d of %d values differ; first at %s: gpu %r oracle %r" % (
        what, int(ne.sum()), ne.size, np.argwhere(ne)[0], got[tuple(np.argwhere(ne)[0])],
        want[tuple(np.argwhere(ne)[0])])


def test_golden_cases_match_reference_outputs(capi, gpu):
    """The committed outputs of the reference's own code, reproduced by the CUDA kernel."""
    gold = np.load(os.path.join(ROOT, "tests", "golden", "refpc_cloud.npz"))
    for c in cloud_cases():
        cam = capi.CameraPair.of(c["P1"], c["P2"])
        if c["mode"] == 0:
            got = gpu.cloud_full(cam, c["dX"], c["dY"], c["img"], bgr=True, sampling=c["sampling"])
        else:
            got = gpu.cloud_foveated(cam, c["dX"], c["dY"], c["fov_h"], c["src_level"], c["img"], bgr=True,
                                     sampling=c["sampling"])
        assert np.array_equal(bits(got), gold[c["name"]]), c["name"]


@pytest.mark.parametrize("shape,sampling,stride,bgr", [((45, 61), 1, 4, True), ((64, 96), 2, 8, True),
                                                       ((33, 130), 3, 4, False), ((1, 1), 1, 4, True),
                                                       ((70, 33), 5, 8, False), ((257, 511), 1, 4, True)])
def test_full_cloud_vs_oracle(capi, gpu, oracle, shape, sampling, stride, bgr):
    rows, cols = shape
    rng = np.random.default_rng(rows * 1000 + cols)
    P1, P2 = cameras(rng, max(cols, 8), max(rows, 8))
    dX = (rng.normal(0, 30, shape) * (rng.random(shape) < 0.9)).astype(np.float32)
    dY = rng.normal(0, 2, shape).astype(np.float32)
    img = rng.integers(0, 256, (rows, cols, 3), dtype=np.uint8)
    got = gpu.cloud_full(capi.CameraPair.of(P1, P2), dX, dY, img, bgr=bgr, sampling=sampling, stride=stride)
    want = oracle.cloud_full(P1, P2, dX, dY, img, bgr=bgr, sampling=sampling, stride=stride)
    same(got, want, "full %s s=%d stride=%d" % (shape, sampling, stride))


@pytest.mark.parametrize("im_shape,level,sampling,stride", [((500, 700), 0, 1, 4), ((500, 700), 3, 1, 8),
                                                            ((421, 633), 6, 2, 4), ((480, 640), 5, 1, 4),
                                                            ((816, 1232), 2, 3, 4)])
def test_foveated_cloud_vs_oracle(capi, gpu, oracle, im_shape, level, sampling, stride):
    im_rows, im_cols = im_shape
    fh, fw = fovea_dims(im_rows, im_cols)
    rng = np.random.default_rng(im_rows + level)
    P1, P2 = cameras(rng, im_cols, im_rows)
    sX = rng.normal(0, 5, (7 * fh, fw)).astype(np.float32)
    sY = rng.normal(0, 1.5, (7 * fh, fw)).astype(np.float32)
    img = rng.integers(0, 256, (im_rows, im_cols, 3), dtype=np.uint8)
    got = gpu.cloud_foveated(capi.CameraPair.of(P1, P2), sX, sY, fh, level, img, sampling=sampling, stride=stride)
    want = oracle.cloud_foveated(P1, P2, sX, sY, fh, level, im_rows, im_cols, img, sampling=sampling, stride=stride)
    same(got, want, "fov %s level %d" % (im_shape, level))


def test_degenerate_geometry_matches_oracle_bits(capi, gpu, oracle):
    """Zero disparity with identical cameras: divisor 0 -> inf / NaN results.  Infinities must agree;
    NaN positions must agree (NaN payload bits are platform noise)."""
    rows, cols = 16, 40
    P = [800.0, 0, 20, 0, 0, 800.0, 8, 0, 0, 0, 1, 0]
    z = np.zeros((rows, cols), np.float32)
    img = np.full((rows, cols, 3), 7, np.uint8)
    got = gpu.cloud_full(capi.CameraPair.of(P, P), z, z, img)
    want = oracle.cloud_full(P, P, z, z, img)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    m = ~np.isnan(want)
    assert np.array_equal(bits(got)[m], bits(want)[m])


def test_cloud_from_full_match_uses_resident_planes(capi, oracle):
    """match_full then cloud_from_match: same points as feeding the returned H, V and the RGB left image
    back through the host entry point and through the oracle; no plane leaves the device in between."""
    L, R, _, _ = synth.make_pair(320, 240, 2)
    rng = np.random.default_rng(3)
    P1, P2 = cameras(rng, 320, 240)
    cam = capi.CameraPair.of(P1, P2)
    with capi.Handle(levels=6) as h:
        H, V, _ = h.match_full(L, R)
        got = h.cloud_from_match(cam, 240, 320, sampling=2)
        assert h.cloud_last_ms() > 0
        again = h.cloud_full(cam, H, V, L, bgr=False, sampling=2)
    want = oracle.cloud_full(P1, P2, H, V, L, bgr=False, sampling=2)
    same(got, want, "from_match full")
    same(again, want, "host entry")


def test_cloud_from_foveated_and_composite_match(capi, oracle):
    L, R, _, _ = synth.make_pair(640, 480, 5)
    rng = np.random.default_rng(4)
    P1, P2 = cameras(rng, 640, 480)
    cam = capi.CameraPair.of(P1, P2)
    with capi.Handle(fovea_level=4) as h:
        sH, sV, _, fw, fh = h.match_foveated(L, R)
        for level in (0, 2, 3):
            got = h.cloud_from_match(cam, fh, fw, src_level=level)
            want = oracle.cloud_foveated(P1, P2, sH, sV, fh, level, 480, 640, L, fovea_level=4, bgr=False)
            same(got, want, "from_match foveated level %d" % level)
        with pytest.raises(capi.UgsmError):
            h.cloud_from_match(cam, 480, 640, src_level=-1)  # foveated match: no full-resolution planes
        H, V, _ = h.match_composite(L, R)
        got = h.cloud_from_match(cam, 480, 640, src_level=-1, sampling=3, stride=8)
        want = oracle.cloud_full(P1, P2, H, V, L, bgr=False, sampling=3, stride=8)
        same(got, want, "from_match composite")


def test_cloud_device_entry_point(capi, oracle):
    import torch

    rows, cols, pitch = 100, 150, 160
    rng = np.random.default_rng(9)
    P1, P2 = cameras(rng, cols, rows)
    dX = rng.normal(0, 10, (rows, cols)).astype(np.float32)
    dY = rng.normal(0, 1, (rows, cols)).astype(np.float32)
    img = rng.integers(0, 256, (rows, cols, 3), dtype=np.uint8)
    tX = torch.zeros((rows, pitch), dtype=torch.float32, device="cuda")
    tY = torch.zeros_like(tX)
    tX[:, :cols] = torch.from_numpy(dX).cuda()
    tY[:, :cols] = torch.from_numpy(dY).cuda()
    tI = torch.from_numpy(img).cuda()
    n = capi.cloud_count(rows, cols, 1)
    tP = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    with capi.Handle() as h:
        got_n = h.cloud_full_device(capi.CameraPair.of(P1, P2), tX.data_ptr(), tY.data_ptr(), pitch, rows, cols,
                                    tI.data_ptr(), cols * 3, True, 1, tP.data_ptr())
    assert got_n == n
    same(tP.cpu().numpy(), oracle.cloud_full(P1, P2, dX, dY, img), "device entry")


def test_cloud_errors(capi, gpu):
    rng = np.random.default_rng(1)
    P1, P2 = cameras(rng, 32, 32)
    cam = capi.CameraPair.of(P1, P2)
    z = np.zeros((8, 8), np.float32)
    img = np.zeros((8, 8, 3), np.uint8)
    with pytest.raises(capi.UgsmError):
        gpu.cloud_full(cam, z, z, img, stride=5)
    with pytest.raises(capi.UgsmError):
        gpu.cloud_full(cam, z, z, img, sampling=0)
    with pytest.raises(capi.UgsmError):
        gpu.cloud_foveated(cam, np.zeros((56, 8), np.float32), np.zeros((56, 8), np.float32), 8, 7, img)
    with capi.Handle() as h, pytest.raises(capi.UgsmError):
        h.cloud_from_match(cam, 8, 8)  # nothing matched yet
    with pytest.raises(capi.UgsmError):
        capi.cloud_count(0, 4, 1)


def test_16mp_full_cloud_vs_oracle(capi, oracle):
    """BASELINE full size: 16,084,992 points from synthetic H, V; every bit against the oracle."""
    rows, cols = 3264, 4928
    rng = np.random.default_rng(16)
    P1, P2 = cameras(rng, cols, rows)
    yy, xx = np.mgrid[0:rows, 0:cols].astype(np.float32)
    dX = (-60 + 18 * np.sin(xx / 301) * np.cos(yy / 517) + rng.normal(0, .2, (rows, cols))).astype(np.float32)
    dY = (1.5 * np.sin(yy / 211) + rng.normal(0, .1, (rows, cols))).astype(np.float32)
    img = rng.integers(0, 256, (rows, cols, 3), dtype=np.uint8)
    with capi.Handle() as h:
        got = h.cloud_full(capi.CameraPair.of(P1, P2), dX, dY, img)
        ms = h.cloud_last_ms()
    import time

    t0 = time.perf_counter()
    want = oracle.cloud_full(P1, P2, dX, dY, img)
    cpu_s = time.perf_counter() - t0
    assert got.shape == (rows * cols, 4)
    same(got, want, "16 MP cloud")
    assert np.isfinite(got[:, :3]).all()
    print("16 MP cloud: kernel %.3f ms on the GPU; oracle port %.3f s on %d host threads" % (
        ms, cpu_s, oracle.num_threads()))


def test_python_mirror_of_the_node_class(capi, gpu, oracle):
    from ug_stereomatcher_b200.pointcloud import CdynamicCalibration

    c = cloud_cases()[1]
    node = CdynamicCalibration(handle=gpu, sampling=c["sampling"])
    with pytest.raises(capi.UgsmError):
        node.doReconstructionRGB(c["dX"], c["dY"])
    assert node.getCameraInfo(c["P1"], c["P2"])
    node.getImages(c["img"])
    same(node.doReconstructionRGB(c["dX"], c["dY"]),
         oracle.cloud_full(c["P1"], c["P2"], c["dX"], c["dY"], c["img"], sampling=c["sampling"]), "mirror full")
    f = cloud_cases()[4]
    node = CdynamicCalibration(handle=gpu, sampling=f["sampling"], foveated=1, srcLevel=f["src_level"])
    node.getCameraInfo(f["P1"], f["P2"])
    node.getImages(f["img"])
    same(node.doReconstructionRGB_FOV(f["dX"], f["dY"]),
         oracle.cloud_foveated(f["P1"], f["P2"], f["dX"], f["dY"], f["fov_h"], f["src_level"], f["img"].shape[0],
                               f["img"].shape[1], f["img"], sampling=f["sampling"]), "mirror foveated")


def test_range_maps_vs_oracle_and_reference_golden(capi, gpu, oracle):
    """Range maps (getRangePoint / getRangePointFOV): K5's range variant against the committed outputs of the unmodified
    reference source and against the oracle, bit for bit; and from the planes a match left on the device."""
    gold = np.load(os.path.join(ROOT, "tests", "golden", "refpc_range.npz"))
    for c in cloud_cases():
        cam = capi.CameraPair.of(c["P1"], c["P2"])
        if c["mode"] == 0:
            got = gpu.range_full(cam, c["dX"], c["dY"])
        else:
            got = gpu.range_foveated(cam, c["dX"], c["dY"], c["fov_h"], c["src_level"], c["img"].shape[0], c["img"].shape[1])
        assert np.array_equal(bits(got).ravel(), gold[c["name"]]), c["name"]
    rng = np.random.default_rng(5)
    rows, cols = 257, 511
    P1, P2 = cameras(rng, cols, rows)
    dX = rng.normal(0, 9, (rows, cols)).astype(np.float32)
    dY = rng.normal(0, 2, (rows, cols)).astype(np.float32)
    same(gpu.range_full(capi.CameraPair.of(P1, P2), dX, dY), oracle.range_full(P1, P2, dX, dY), "range 257x511")
    from ug_stereomatcher_b200 import synth
    L, R, _, _ = synth.make_pair(320, 240, 3)
    P1, P2 = cameras(rng, 320, 240)
    with capi.Handle(levels=5) as h:
        H, V, _ = h.match_full(L, R)
        z = h.range_from_match(capi.CameraPair.of(P1, P2), 240, 320)
    same(z, oracle.range_full(P1, P2, H, V), "range from the resident planes")
