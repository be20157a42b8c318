"""-m gpu: the reference's own matcher node source (src/gpu_matcher/UG_GPU_matcher.cpp), compiled unmodified
against include/MatchGPULib.h through inert ROS / OpenCV / cv_bridge stand-ins (oracle/_ref/refnode, built by
`make -C oracle` wherever /root/reference is mounted), run on libugsm.so: what it publishes must be bit-equal
to the C-ABI results.  This is the drop-in claim of INTEGRATION.md section 1, executed."""
import os
import struct
import subprocess

import numpy as np
import pytest

from ug_stereomatcher_b200 import synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "refnode")
TOPICS = ["output_disparityH", "output_disparityV", "output_disparityC", "output_stackH", "output_stackV",
          "output_stackC", "output_stackL_pyramid", "output_stackR_pyramid"]
SRV_FIELDS = ["srv_dispH", "srv_dispV", "srv_dispC", "srv_fdispH", "srv_fdispV", "srv_fdispC"]


def run_node(tmp_path, L, R, foveated, node_args=(), service=False):
    rows, cols, _ = L.shape
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("3i", foveated | (2 if service else 0), rows, cols))
        f.write(np.ascontiguousarray(L).tobytes())
        f.write(np.ascontiguousarray(R).tobytes())
    p = subprocess.run([EXE, fin, fout, *node_args], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    raw = open(fout, "rb").read()
    out, off = {}, 0
    for t in TOPICS + SRV_FIELDS:
        r, c = struct.unpack_from("2i", raw, off)
        off += 8
        if r > 0:
            out[t] = np.frombuffer(raw, np.float32, r * c, off).reshape(r, c).copy()
            off += 4 * r * c
    return out


needs_node = pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref/refnode not built (needs /root/reference)")


@needs_node
def test_reference_node_publishes_cabi_result_full(capi, tmp_path):
    L, R, _, _ = synth.make_pair(640, 480, 9)
    pub = run_node(tmp_path, L, R, foveated=0)
    assert sorted(pub) == ["output_disparityC", "output_disparityH", "output_disparityV"]
    with capi.Handle() as h:
        H, V, C = h.match_full(L, R)
    assert np.array_equal(pub["output_disparityH"], H)
    assert np.array_equal(pub["output_disparityV"], V)
    assert np.array_equal(pub["output_disparityC"], C)


@needs_node
@pytest.mark.parametrize("fovea_level", [7, 4])
def test_reference_node_publishes_cabi_result_foveated(capi, tmp_path, fovea_level):
    """foveated = 1: initStack + matchStackPyramid (UG_GPU_matcher.cpp:160-303); the node's argv[2] is the fovea
    level.  Stacks are level-major; the image pyramid stacks carry 3 channels per level (rows k*3*fovH + ch*fovH + i)."""
    L, R, _, _ = synth.make_pair(640, 480, 10)
    pub = run_node(tmp_path, L, R, foveated=1, node_args=("x", str(fovea_level)))
    assert sorted(pub) == sorted(TOPICS[3:])
    with capi.Handle(fovea_level=fovea_level) as h:
        sH, sV, sC, fw, fh = h.match_foveated(L, R)
        assert np.array_equal(pub["output_stackH"], sH)
        assert np.array_equal(pub["output_stackV"], sV)
        assert np.array_equal(pub["output_stackC"], sC)
        assert pub["output_stackL_pyramid"].shape == (fovea_level * fh * 3, fw)
        for k in range(fovea_level):
            Lk, Rk = h.fovea_images(k, 480, 640)
            if Lk.shape[1:] != (fh, fw):
                continue  # level fovea_level-1 .. keep to the fovea-sized levels the node indexes with fovW x fovH
            for ch in range(3):
                r0 = k * fh * 3 + ch * fh
                assert np.array_equal(pub["output_stackL_pyramid"][r0:r0 + fh], Lk[ch]), (k, ch)
                assert np.array_equal(pub["output_stackR_pyramid"][r0:r0 + fh], Rk[ch]), (k, ch)


@needs_node
def test_reference_node_service_full(capi, tmp_path):
    """disparitySrv (UG_GPU_matcher.cpp:497-690), foveated = 0: the GetDisparitiesGPU response carries dispH / dispV /
    dispC .image as 32FC1 planes (UG_GPU_matcher.cpp:648-683) -- bit-equal to the C-ABI result; nothing is published."""
    L, R, _, _ = synth.make_pair(640, 480, 11)
    out = run_node(tmp_path, L, R, foveated=0, service=True)
    assert sorted(out) == ["srv_dispC", "srv_dispH", "srv_dispV"]
    with capi.Handle() as h:
        H, V, C = h.match_full(L, R)
    assert np.array_equal(out["srv_dispH"], H)
    assert np.array_equal(out["srv_dispV"], V)
    assert np.array_equal(out["srv_dispC"], C)


@needs_node
def test_reference_node_service_foveated_keeps_the_reference_copy_loop(capi, tmp_path):
    """disparitySrv, foveated = 1 (matchStack): the node source is the reference's, so its response is what the
    reference's copy loops make of the stack -- UG_GPU_matcher.cpp:559-563 reads plane k with fovH as the row pitch
    and :572-578 fills only the first fovH rows of the (levels * fovH) x fovW image.  The planes it
    copies from are this library's: the response equals those loops applied to the C-ABI stack, bit for bit.
    (INTEGRATION.md shows the two-line correction for maintainers; the topic path and the C-ABI are not affected.)"""
    L, R, _, _ = synth.make_pair(640, 480, 12)
    out = run_node(tmp_path, L, R, foveated=1, node_args=("x", "7"), service=True)
    assert sorted(out) == ["srv_fdispC", "srv_fdispH", "srv_fdispV"]
    with capi.Handle(fovea_level=7) as h:
        sH, sV, sC, fw, fh = h.match_foveated(L, R)
    for name, stack in (("srv_fdispH", sH), ("srv_fdispV", sV), ("srv_fdispC", sC)):
        got = out[name]
        assert got.shape == (7 * fh, fw)
        level0 = stack[:fh].ravel()  # plane of level 0, fovW x fovH floats
        want = np.zeros_like(got)
        ii, jj = np.mgrid[0:fh, 0:fw]
        want[:fh] = level0[ii * fh + jj]
        assert np.array_equal(got, want), name
