"""Real-image feeder (include/ugsm.h ugsm_feeder_* / ugsm_tiff_*): CPU tests.

The decoder is pinned against cv::imread(path, 1) -- the call the reference's camera simulator makes
(include/ug_stereomatcher/settings.h:42) -- through fixtures generated with cv2 in the build container
(tests/golden/make_tiff_golden.py) and, when cv2 is importable, live on freshly written files.  The list
semantics follow Settings::readStringList / nextImage (settings.h:15-68).  Host-only: no GPU needed."""
import os
import threading

import numpy as np
import pytest

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiff")
with np.load(os.path.join(HERE, "expected_bgr.npz")) as _z:
    EXPECTED = {k: _z[k] for k in _z.files}  # eager: NpzFile reads lazily and is not thread-safe


@pytest.mark.parametrize("name", sorted(EXPECTED))
def test_tiff_decode_matches_cv_imread_golden(capi, name):
    got = capi.tiff_read(os.path.join(HERE, name))
    assert got.dtype == np.uint8 and got.shape == EXPECTED[name].shape
    assert np.array_equal(got, EXPECTED[name])  # bit-exact: byte data


def test_tiff_decode_live_against_cv2(capi, tmp_path):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    for k, (h, w) in enumerate([(1, 1), (7, 300), (131, 97), (260, 515)]):
        a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        a[: h // 2] = a[: h // 2] // 64 * 64  # long runs: exercises the LZW table growth and resets
        for tag, params in (("lzw", []), ("none", [cv2.IMWRITE_TIFF_COMPRESSION, 1]),
                            ("packbits", [cv2.IMWRITE_TIFF_COMPRESSION, 32773]),
                            ("deflate", [cv2.IMWRITE_TIFF_COMPRESSION, 8])):
            p = str(tmp_path / ("%d_%s.tif" % (k, tag)))
            assert cv2.imwrite(p, a, params)
            assert np.array_equal(capi.tiff_read(p), cv2.imread(p, 1)), (h, w, tag)
        a16 = rng.integers(0, 65536, (h, w, 3), dtype=np.uint16)
        p = str(tmp_path / ("%d_16.tif" % k))
        assert cv2.imwrite(p, a16)
        assert np.array_equal(capi.tiff_read(p), cv2.imread(p, 1)), (h, w, "rgb16")
    big = rng.integers(0, 4, (700, 900, 3), dtype=np.uint8) * 60  # > 4096 LZW codes per strip: several clear codes
    p = str(tmp_path / "big.tif")
    assert cv2.imwrite(p, big)
    assert np.array_equal(capi.tiff_read(p), cv2.imread(p, 1))


def test_tiff_errors_are_statuses_not_crashes(capi, tmp_path):
    p = tmp_path / "junk.tif"
    p.write_bytes(b"II*\0" + b"\xff" * 40)
    with pytest.raises(capi.UgsmError):
        capi.tiff_read(str(p))
    with pytest.raises(capi.UgsmError) as e:
        capi.tiff_read(str(tmp_path / "missing.tif"))
    assert "cannot open" in str(e.value)
    good = open(os.path.join(HERE, "rgb_lzw.tif"), "rb").read()
    t = tmp_path / "trunc.tif"
    t.write_bytes(good[: len(good) // 2])
    with pytest.raises(capi.UgsmError):
        capi.tiff_read(str(t))


def test_feeder_follows_the_list_in_pairs(capi):
    names = ["rgb_none.tif", "cv_rgb8.tif", "rgb_lzw.tif", "raw_tiles.tif", "rgb_packbits.tif", "rgb_deflate.tif"]
    with capi.Feeder(os.path.join(HERE, "input_images.xml"), slots=2, threads=2) as f:
        fi = f.info()
        assert (fi.rows, fi.cols, fi.step, fi.list_images, fi.pairs, fi.slots) == (45, 61, 183, 6, 3, 2)
        for i in range(3):
            p, left, right = f.next()
            assert p.index == i and (p.rows, p.cols) == (45, 61)
            assert np.array_equal(left, EXPECTED[names[2 * i]])
            assert np.array_equal(right, EXPECTED[names[2 * i + 1]])
            f.release(p.slot)
        assert f.next() is None and f.next() is None  # UGSM_FEEDER_END, repeatedly


def test_feeder_past_the_end_repeats_the_last_pair(capi):
    """Settings::nextImage: `if (atImageList >= size) atImageList = size - 2` (settings.h:34-38)."""
    with capi.Feeder(os.path.join(HERE, "input_images.xml"), pairs=5, slots=3, threads=3) as f:
        seen = []
        for i in range(5):
            p, left, right = f.next()
            seen.append((left.copy(), right.copy()))
            f.release(p.slot)
        assert f.next() is None
    for k in (3, 4):
        assert np.array_equal(seen[k][0], EXPECTED["rgb_packbits.tif"])
        assert np.array_equal(seen[k][1], EXPECTED["rgb_deflate.tif"])


def test_feeder_ring_blocks_until_release_and_serves_concurrent_consumers(capi):
    with capi.Feeder(os.path.join(HERE, "input_images.xml"), pairs=40, slots=3, threads=4) as f:
        got, lock = [], threading.Lock()

        def consume():
            while True:
                r = f.next()
                if r is None:
                    return
                p, left, right = r
                ok = np.array_equal(left, EXPECTED["rgb_none.tif" if p.index == 0 else
                                                   "rgb_lzw.tif" if p.index == 1 else "rgb_packbits.tif"])
                with lock:
                    got.append((p.index, ok))
                f.release(p.slot)

        ts = [threading.Thread(target=consume) for _ in range(3)]
        [t.start() for t in ts]
        [t.join(60) for t in ts]
        assert sorted(i for i, _ in got) == list(range(40)) and all(ok for _, ok in got)
        with pytest.raises(capi.UgsmError):
            f.release(0)  # nothing handed out any more


def test_feeder_rejects_bad_lists(capi, tmp_path):
    bad = tmp_path / "l.xml"
    bad.write_text("<?xml version=\"1.0\"?><opencv_storage><images>only_one.tif</images></opencv_storage>")
    with pytest.raises(capi.UgsmError):
        capi.Feeder(str(bad))
    bad.write_text("not xml at all")
    with pytest.raises(capi.UgsmError):
        capi.Feeder(str(bad))
    # a pair whose image has another size fails for that pair only
    import shutil
    for n in ("rgb_none.tif", "rgb_lzw.tif"):
        shutil.copy(os.path.join(HERE, n), tmp_path / n)
    cvw = pytest.importorskip("cv2")
    cvw.imwrite(str(tmp_path / "small.tif"), np.zeros((5, 5, 3), np.uint8))
    bad.write_text("<?xml version=\"1.0\"?><opencv_storage><images>rgb_none.tif small.tif rgb_none.tif rgb_lzw.tif"
                   "</images></opencv_storage>")
    with capi.Feeder(str(bad), slots=2, threads=1) as f:
        with pytest.raises(capi.UgsmError) as e:
            f.next()
        assert "size" in str(e.value)
        p, left, right = f.next()
        assert p.index == 1 and np.array_equal(right, EXPECTED["rgb_lzw.tif"])
        f.release(p.slot)


def write_tiff_rgb(path, rgb):
    """Minimal baseline TIFF (little-endian, one uncompressed strip, 8-bit RGB) -- no imaging library needed."""
    import struct
    h, w, _ = rgb.shape
    tags = [(256, 3, w), (257, 3, h), (259, 3, 1), (262, 3, 2), (277, 3, 3), (278, 3, h), (284, 3, 1)]
    n = len(tags) + 3
    bps_off = 8 + 2 + 12 * n + 4
    data_off = bps_off + 6
    tags += [(258, 3, None), (273, 4, data_off), (279, 4, h * w * 3)]
    tags.sort()
    ifd = struct.pack("<H", n)
    for tag, typ, v in tags:
        if v is None:
            ifd += struct.pack("<HHII", tag, 3, 3, bps_off)
        elif typ == 3:
            ifd += struct.pack("<HHIHH", tag, 3, 1, v, 0)
        else:
            ifd += struct.pack("<HHII", tag, 4, 1, v)
    blob = b"II" + struct.pack("<HI", 42, 8) + ifd + struct.pack("<I", 0) + struct.pack("<HHH", 8, 8, 8)
    with open(path, "wb") as f:
        f.write(blob + np.ascontiguousarray(rgb, np.uint8).tobytes())


def test_minimal_tiff_writer_round_trips(capi, tmp_path):
    a = np.random.default_rng(0).integers(0, 256, (33, 47, 3), dtype=np.uint8)
    write_tiff_rgb(str(tmp_path / "a.tif"), a)
    assert np.array_equal(capi.tiff_read(str(tmp_path / "a.tif")), a[:, :, ::-1])


@pytest.mark.gpu
def test_batch_driver_matches_a_tiff_list(capi, tmp_path):
    """ugsm_batch --list: TIFF pairs -> feeder ring (page-locked) -> ugsm_match_full; the last pair's result equals the
    single C-ABI call on the same bytes (pairs beyond the list repeat the last pair, settings.h:34-38)."""
    import json
    import subprocess

    from ug_stereomatcher_b200 import synth

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    batch = os.path.join(root, "ug_stereomatcher_b200", "ugsm_batch")
    pairs = [synth.make_pair(640, 480, s)[:2] for s in (1, 2)]
    names = []
    for k, (L, R) in enumerate(pairs):
        for tag, img in (("l", L), ("r", R)):
            write_tiff_rgb(str(tmp_path / ("%s%d.tif" % (tag, k))), img[:, :, ::-1])  # decoded bytes == img
            names.append("%s%d.tif" % (tag, k))
    (tmp_path / "list.xml").write_text('<?xml version="1.0"?>\n<opencv_storage>\n<images>\n%s\n</images>\n'
                                       "</opencv_storage>\n" % "\n".join(names))
    p = subprocess.run([batch, "0", "0", "5", "--in-flight", "1", "--list", str(tmp_path / "list.xml"),
                        "--decode-threads", "2"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode == 0, p.stderr.decode()
    rec = json.loads(p.stdout.decode())
    assert rec["pairs"] == 5 and rec["rows"] == 480 and rec["cols"] == 640 and rec["input"] == "tiff list"
    assert rec["list_images"] == 4 and rec["ring_pinned"] == 1
    with capi.Handle() as h:
        want = h.match_full(*pairs[1])[0]
    assert rec["h_sum"] == pytest.approx(float(want.astype(np.float64).sum()), rel=1e-9, abs=1e-6)


def test_feeder_odd_list_follows_the_reference_reader(capi, tmp_path):
    """Settings::nextImage with an odd number of entries: after the last entry the index is reset to size - 2, so the
    second pair of a 3-entry list is (entry 2, entry 1), then (2, 1) again."""
    import shutil
    names = ["rgb_none.tif", "rgb_lzw.tif", "rgb_packbits.tif"]
    for n in names:
        shutil.copy(os.path.join(HERE, n), tmp_path / n)
    (tmp_path / "l.xml").write_text('<?xml version="1.0"?>\n<opencv_storage>\n<images>\n"%s"\n%s\n%s\n</images>\n'
                                    "</opencv_storage>\n" % tuple(names))  # a quoted entry as FileStorage allows
    with capi.Feeder(str(tmp_path / "l.xml"), pairs=3, slots=2, threads=2) as f:
        assert f.info().list_images == 3
        want = [(0, 1), (2, 1), (2, 1)]
        for i, (a, b) in enumerate(want):
            p, left, right = f.next()
            assert p.index == i
            assert np.array_equal(left, EXPECTED[names[a]]) and np.array_equal(right, EXPECTED[names[b]])
            f.release(p.slot)
        assert f.next() is None


def test_tiff_decoder_survives_corrupted_files(capi, tmp_path):
    """Flipped bytes and truncation in every fixture: the decoder either returns pixels or a status, never crashes
    or hangs (the same corpus ran clean under -fsanitize=address,undefined during development)."""
    rng = np.random.default_rng(5)
    names = sorted(n for n in os.listdir(HERE) if n.endswith(".tif"))
    p = str(tmp_path / "fz.tif")
    decoded = rejected = 0
    for it in range(400):
        b = bytearray(open(os.path.join(HERE, names[it % len(names)]), "rb").read())
        for _ in range(int(rng.integers(1, 12))):
            b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        if it % 7 == 0:
            b = b[: int(rng.integers(8, len(b)))]
        with open(p, "wb") as f:
            f.write(bytes(b))
        try:
            a = capi.tiff_read(p)
            assert a.dtype == np.uint8 and a.ndim == 3
            decoded += 1
        except capi.UgsmError:
            rejected += 1
    assert decoded + rejected == 400 and rejected > 0
