#!/usr/bin/env python
"""Generate tests/golden/refpc_cloud.npz: point clouds produced by the UNMODIFIED reference point-cloud
source (src/pointcloud/getPointCloud.cpp) compiled through oracle/refbuild/stubs_pc into oracle/_ref/refpc
(`make -C oracle`, CPU only, needs /root/reference).  The inputs are regenerated deterministically by
cloud_cases() below (shared with tests/test_cloud_oracle.py); only the outputs are stored.

Run here:   python tests/golden/make_refpc_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def cameras(rng, w, h):
    """A plausible verged pair: left K[I|0], right with slightly different intrinsics, a baseline and a
    small projective row, so that every entry of P2 the formula reads is non-trivial."""
    f = rng.uniform(0.6, 1.4) * w
    P1 = np.array([[f, 0, w / 2 + rng.normal(0, 5), 0], [0, f * rng.uniform(.99, 1.01), h / 2 + rng.normal(0, 5), 0],
                   [0, 0, 1, 0]])
    P2 = np.array([[f * rng.uniform(.98, 1.02), rng.normal(0, 2), w / 2 + rng.normal(0, 9), -f * rng.uniform(50, 300)],
                   [rng.normal(0, 2), f * rng.uniform(.98, 1.02), h / 2 + rng.normal(0, 9), rng.normal(0, 500)],
                   [rng.normal(0, 1e-5), rng.normal(0, 1e-5), 1, rng.normal(0, .3)]])
    return P1, P2


def fovea_dims(im_rows, im_cols, fovea_level=7):
    w, h = im_cols, im_rows
    for _ in range(fovea_level - 1):
        w, h = int(w / 1.41421356), int(h / 1.41421356)
    return h, w


def cloud_cases():
    """-> list of dicts: name, mode (0 full / 1 foveated), P1, P2, dX, dY, img, sampling, src_level, fov_h."""
    rng = np.random.default_rng(0xC10D)
    cases = []
    for k, (rows, cols, s) in enumerate([(45, 61, 1), (64, 96, 2), (33, 130, 3)]):
        P1, P2 = cameras(rng, cols, rows)
        cases.append(dict(name="full%d" % k, mode=0, P1=P1, P2=P2,
                          dX=rng.normal(0, 8, (rows, cols)).astype(np.float32),
                          dY=rng.normal(0, 1.5, (rows, cols)).astype(np.float32),
                          img=rng.integers(0, 256, (rows, cols, 3), dtype=np.uint8), sampling=s, src_level=0,
                          fov_h=0))
    for k, (im_rows, im_cols, s, lev) in enumerate([(500, 700, 1, 0), (500, 700, 1, 3), (421, 633, 2, 6),
                                                    (480, 640, 1, 5)]):
        fh, fw = fovea_dims(im_rows, im_cols)
        P1, P2 = cameras(rng, im_cols, im_rows)
        cases.append(dict(name="fov%d" % k, mode=1, P1=P1, P2=P2,
                          dX=rng.normal(0, 3, (7 * fh, fw)).astype(np.float32),
                          dY=rng.normal(0, 1, (7 * fh, fw)).astype(np.float32),
                          img=rng.integers(0, 256, (im_rows, im_cols, 3), dtype=np.uint8), sampling=s,
                          src_level=lev, fov_h=fh))
    return cases


def main():
    import ugsm_oracle as o

    out = {}
    for c in cloud_cases():
        pts = o.run_refpc(c["mode"], c["P1"], c["P2"], c["dX"], c["dY"], c["img"], sampling=c["sampling"],
                          src_level=c["src_level"])
        out[c["name"]] = pts.view(np.uint32)  # raw bits: exact comparison, NaN-safe
        print(c["name"], pts.shape)
    path = os.path.join(ROOT, "tests", "golden", "refpc_cloud.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
    # range maps of the same inputs: getRangePoint / getRangePointFOV at every pixel (refpc modes 2 / 3)
    rng_out = {}
    for c in cloud_cases():
        r = o.run_refpc(2 + c["mode"], c["P1"], c["P2"], c["dX"], c["dY"], c["img"], src_level=c["src_level"])
        rng_out[c["name"]] = r.view(np.uint32)
    path = os.path.join(ROOT, "tests", "golden", "refpc_range.npz")
    np.savez_compressed(path, **rng_out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
