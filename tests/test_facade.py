"""The reference-named host side: C++ MatchGPULib facade / headless driver and its Python mirror."""
import os
import subprocess

import numpy as np
import pytest

from ug_stereomatcher_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "ug_stereomatcher_b200", "ugsm_headless")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_facade_symbols_exported(capi):
    out = subprocess.run(["nm", "-DC", capi.LIB_PATH], stdout=subprocess.PIPE, check=True).stdout.decode()
    for name in ["MatchGPULib::MatchGPULib(int, char**)", "MatchGPULib::match(", "MatchGPULib::initStack(",
                 "MatchGPULib::matchStack(", "MatchGPULib::matchStackPyramid(", "MatchGPULib::getFoveaWidth()",
                 "MatchGPULib::getFoveaHeight()", "MatchGPULib::getFoveateLevel()", "MatchGPULib::setFoveated(int)",
                 "MatchGPULib::setFoveaWidth(int)", "MatchGPULib::setFoveaHeight(int)", "ugsmFreeStack("]:
        assert name in out, name


def test_headless_fails_loudly_without_gpu(capi, tmp_path):
    if _have_gpu():
        pytest.skip("GPU present")
    assert os.path.exists(EXE)
    np.zeros((8, 8, 3), np.uint8).tofile(tmp_path / "z.rgb")
    p = subprocess.run([EXE, "8", "8", str(tmp_path / "z.rgb"), str(tmp_path / "z.rgb"), str(tmp_path / "o.f32")],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode == 3 and b"no CPU fallback" in p.stderr


@pytest.mark.gpu
def test_headless_full_matches_cabi(capi, tmp_path):
    L, R, _, _ = synth.make_pair(320, 240, 0)
    L.tofile(tmp_path / "l.rgb")
    R.tofile(tmp_path / "r.rgb")
    p = subprocess.run([EXE, "240", "320", str(tmp_path / "l.rgb"), str(tmp_path / "r.rgb"), str(tmp_path / "o.f32"),
                        "--repeat", "2"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode == 0, p.stderr.decode()
    out = np.fromfile(tmp_path / "o.f32", np.float32).reshape(3, 240, 320)
    with capi.Handle() as h:
        want = h.match_full(L, R)
    for k in range(3):
        assert (out[k] == want[k]).all()


@pytest.mark.gpu
def test_headless_foveated_matches_cabi(capi, tmp_path):
    L, R, _, _ = synth.make_pair(1232, 816, 1)
    L.tofile(tmp_path / "l.rgb")
    R.tofile(tmp_path / "r.rgb")
    p = subprocess.run([EXE, "816", "1232", str(tmp_path / "l.rgb"), str(tmp_path / "r.rgb"), str(tmp_path / "o.f32"),
                        "--foveated"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode == 0, p.stderr.decode()
    with capi.Handle() as h:
        sH, sV, sC, fw, fh = h.match_foveated(L, R)
    out = np.fromfile(tmp_path / "o.f32", np.float32).reshape(3, 7 * fh, fw)
    assert (out[0] == sH).all() and (out[1] == sV).all() and (out[2] == sC).all()


@pytest.mark.gpu
def test_python_mirror(capi, oracle):
    from ug_stereomatcher_b200 import MatchGPULib

    L, R, _, _ = synth.make_pair(1232, 816, 2)
    m = MatchGPULib(["node", "x", "7"])
    m.setFoveated(1)
    m.initStack(L, R)
    w, h = oracle.level_dims(1232, 816)
    assert (m.getFoveaWidth(), m.getFoveaHeight(), m.getFoveateLevel()) == (w[6], h[6], 7)
    disp, lf, rf = m.matchStackPyramid(L, R)
    assert len(disp) == 14 and disp[0][0].shape == (h[6], w[6]) and disp[9][0].shape == (h[9], w[9])
    # image foveas: centre crops of the pyramid (MatchGPULib.cpp:1171-1185)
    pyr = oracle.pyramid_rgb(L, 14)
    l, u = w[2] // 2 - w[6] // 2, h[2] // 2 - h[6] // 2
    assert (lf[2] == pyr[2][:, u:u + h[6], l:l + w[6]]).all()
    assert (lf[8] == pyr[8]).all()
    H, V, C = m.match(L, R, 0)
    assert H.shape == (816, 1232)
    Hc, Vc, Cc = m.match(L, R, 1)  # foveated + hierarchicalDisparity composite
    want = oracle.match_composite(L, R)
    assert (Hc == want[0]).all() and (Vc == want[1]).all() and (Cc == want[2]).all()
    m.close()


@pytest.mark.skipif(not os.path.exists("/root/reference/src/gpu_matcher/UG_GPU_matcher.cpp"),
                    reason="needs the reference checkout")
def test_reference_node_source_builds_against_the_facade(capi, tmp_path):
    """The reference's matcher node (UG_GPU_matcher.cpp), unmodified, compiles against include/MatchGPULib.h and
    links with libugsm.so (ROS / OpenCV / cv_bridge replaced by inert stand-ins).  CPU-only: compile + link."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "refnode")
    cmd = ["/usr/bin/g++", "-O0", "-std=c++14", "-fpermissive", "-w", "-DUGSM_WITH_CV_BRIDGE",
           "-I" + os.path.join(root, "oracle", "refbuild", "stubs_pc"), "-I" + os.path.join(root, "include"),
           "-I/root/reference/include", "-I/root/reference", "-o", exe,
           os.path.join(root, "oracle", "refbuild", "refnode_main.cpp"),
           os.path.join(root, "ug_stereomatcher_b200", "csrc", "MatchGPULib.cpp"),
           "-L" + os.path.join(root, "ug_stereomatcher_b200"), "-lugsm",
           "-Wl,-rpath," + os.path.join(root, "ug_stereomatcher_b200")]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode == 0, p.stderr.decode()[-3000:]
    assert os.path.exists(exe)


BATCH = os.path.join(ROOT, "ug_stereomatcher_b200", "ugsm_batch")


def test_batch_driver_fails_loudly_without_gpu(capi):
    if _have_gpu():
        pytest.skip("GPU present")
    assert os.path.exists(BATCH)
    p = subprocess.run([BATCH, "240", "320", "2"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode == 3 and b"no CPU fallback" in p.stderr and p.stdout == b""


@pytest.mark.gpu
@pytest.mark.parametrize("foveated", [False, True])
def test_batch_driver_streams_pairs_and_matches_cabi(capi, tmp_path, foveated):
    """ugsm_batch: worker threads with one handle + pinned buffers each (ugsm_host_alloc) pull pairs from a shared
    counter; every pair of the batch is the same image pair here, so the last result must equal the single call."""
    import json

    L, R, _, _ = synth.make_pair(640, 480, 3)
    L.tofile(tmp_path / "l.rgb")
    R.tofile(tmp_path / "r.rgb")
    cmd = [BATCH, "480", "640", "7", "--in-flight", "3", "--left", str(tmp_path / "l.rgb"), "--right",
           str(tmp_path / "r.rgb")] + (["--foveated"] if foveated else [])
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode == 0, p.stderr.decode()
    rec = json.loads(p.stdout.decode())
    assert rec["pairs"] == 7 and rec["gpus"] == 1 and rec["in_flight_per_gpu"] == 3 and rec["pairs_per_s"] > 0
    with capi.Handle() as h:
        want = h.match_foveated(L, R)[0] if foveated else h.match_full(L, R)[0]
    assert rec["h_sum"] == pytest.approx(float(want.astype(np.float64).sum()), rel=1e-9, abs=1e-6)
