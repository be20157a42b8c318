"""Writes the small TIFF fixtures of tests/golden/tiff/ and, beside each, the pixels cv2.imread(path, 1) returns
(the decoder the reference's camera simulator calls: cv::imread(.., 1), include/ug_stereomatcher/settings.h:42).
Run where PIL and cv2 are importable (this build container); the fixtures travel, the libraries need not.

    python tests/golden/make_tiff_golden.py
"""
import os

import cv2
import numpy as np
from PIL import Image

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tiff")


def main():
    os.makedirs(HERE, exist_ok=True)
    rng = np.random.default_rng(7)
    h, w = 45, 61
    yy, xx = np.mgrid[0:h, 0:w]
    smooth = np.stack([(xx * 4 + yy) % 256, (yy * 5 + xx // 2) % 256, (xx + yy) * 2 % 256], -1).astype(np.uint8)
    noise = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    a = np.where((xx[..., None] // 8 + yy[..., None] // 8) % 2 == 0, smooth, noise).astype(np.uint8)
    a16 = (a.astype(np.uint16) << 8) | rng.integers(0, 256, (h, w, 3)).astype(np.uint16)
    cases = {}

    def pil(name, arr, **kw):
        Image.fromarray(arr).save(os.path.join(HERE, name), **kw)
        cases[name] = None

    pil("rgb_none.tif", a)
    pil("rgb_lzw.tif", a, compression="tiff_lzw")
    pil("rgb_packbits.tif", a, compression="packbits")
    pil("rgb_deflate.tif", a, compression="tiff_adobe_deflate")
    pil("grey.tif", a[:, :, 1])
    pil("grey_lzw.tif", a[:, :, 1], compression="tiff_lzw")
    Image.fromarray(a).convert("P", palette=Image.ADAPTIVE).save(os.path.join(HERE, "palette.tif"))
    cases["palette.tif"] = None
    # multi-strip, big-endian, planar and tiled variants through PIL's libtiff options
    Image.fromarray(a).save(os.path.join(HERE, "rgb_strips8.tif"), tiffinfo={278: 8})
    cases["rgb_strips8.tif"] = None
    # what OpenCV itself writes (LZW + horizontal predictor, several strips), 8 and 16 bit
    cv2.imwrite(os.path.join(HERE, "cv_rgb8.tif"), a)
    cv2.imwrite(os.path.join(HERE, "cv_rgb16.tif"), a16)
    cv2.imwrite(os.path.join(HERE, "cv_grey16.tif"), a16[:, :, 0])
    for n in ("cv_rgb8.tif", "cv_rgb16.tif", "cv_grey16.tif"):
        cases[n] = None
    # hand-written variants PIL / cv2 do not emit: big-endian, planar configuration 2, tiles
    import struct

    def write_raw(name, arr, be=False, planar=False, tile=None):
        e = ">" if be else "<"
        hh, ww, cc = arr.shape
        chunks = []
        if tile:
            tw, th = tile
            for ty in range(0, hh, th):
                for tx in range(0, ww, tw):
                    t = np.zeros((th, tw, cc), np.uint8)
                    blk = arr[ty:ty + th, tx:tx + tw]
                    t[:blk.shape[0], :blk.shape[1]] = blk
                    chunks.append(t.tobytes())
        elif planar:
            for c in range(cc):
                chunks.append(arr[:, :, c].tobytes())
        else:
            rps = 16
            for y in range(0, hh, rps):
                chunks.append(arr[y:y + rps].tobytes())
        tags = [(256, 3, [ww]), (257, 3, [hh]), (258, 3, [8] * cc), (259, 3, [1]), (262, 3, [2]), (277, 3, [cc]),
                (284, 3, [2 if planar else 1])]
        if tile:
            tags += [(322, 3, [tile[0]]), (323, 3, [tile[1]]), (324, 4, None), (325, 4, [len(c) for c in chunks])]
        else:
            tags += [(273, 4, None), (278, 3, [hh if planar else 16]), (279, 4, [len(c) for c in chunks])]
        tags.sort()
        n = len(tags)
        ifd_off = 8
        data_off = ifd_off + 2 + 12 * n + 4
        extra = b""
        body_entries = []
        # first pass: sizes of out-of-line values, then chunk offsets
        sizes = {3: 2, 4: 4}
        outl = 0
        for tag, typ, vals in tags:
            cnt = len(chunks) if vals is None else len(vals)
            if sizes[typ] * cnt > 4:
                outl += sizes[typ] * cnt + (sizes[typ] * cnt) % 2
        chunk_off = data_off + outl
        offs = []
        for c in chunks:
            offs.append(chunk_off)
            chunk_off += len(c)
        for tag, typ, vals in tags:
            if vals is None:
                vals = offs
            cnt = len(vals)
            fmt = {3: "H", 4: "I"}[typ]
            raw = struct.pack(e + fmt * cnt, *vals)
            if len(raw) > 4:
                body_entries.append(struct.pack(e + "HHII", tag, typ, cnt, data_off + len(extra)))
                extra += raw + (b"\0" if len(raw) % 2 else b"")
            else:
                body_entries.append(struct.pack(e + "HHI", tag, typ, cnt) + raw.ljust(4, b"\0"))
        hdr = (b"MM" if be else b"II") + struct.pack(e + "HI", 42, ifd_off)
        blob = hdr + struct.pack(e + "H", n) + b"".join(body_entries) + struct.pack(e + "I", 0) + extra + b"".join(chunks)
        open(os.path.join(HERE, name), "wb").write(blob)
        cases[name] = None

    write_raw("raw_be.tif", a, be=True)
    write_raw("raw_planar.tif", a, planar=True)
    write_raw("raw_tiles.tif", a, tile=(32, 16))
    write_raw("raw_be_tiles.tif", a, be=True, tile=(16, 16))

    out = {}
    for name in sorted(cases):
        img = cv2.imread(os.path.join(HERE, name), 1)
        assert img is not None and img.dtype == np.uint8 and img.shape == (h, w, 3), name
        out[name] = img
    np.savez_compressed(os.path.join(HERE, "expected_bgr.npz"), **out)
    # the image list the reference's simulator reads (src/utils/input_images.xml), relative paths, 3 pairs
    names = ["rgb_none.tif", "cv_rgb8.tif", "rgb_lzw.tif", "raw_tiles.tif", "rgb_packbits.tif", "rgb_deflate.tif"]
    with open(os.path.join(HERE, "input_images.xml"), "w") as f:
        f.write('<?xml version="1.0"?>\n<opencv_storage>\n<!-- left / right, pair by pair -->\n<images>\n')
        f.write("\n".join(names))
        f.write("\n</images>\n</opencv_storage>\n")
    print("wrote", len(out), "fixtures,", sum(os.path.getsize(os.path.join(HERE, n)) for n in os.listdir(HERE)), "bytes")


if __name__ == "__main__":
    main()
