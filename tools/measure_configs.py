#!/usr/bin/env python
"""Device time of every BASELINE.json config on one B200 (one pair in flight, CUDA events inside the
library), plus a sanity end-point error against the synthetic ground truth.  Prints one JSON line each."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from ug_stereomatcher_b200 import capi, synth


def run(name, w, h, mode, reps=5, **cfg):
    L, R, dH, dV = synth.make_pair(w, h, 0)
    with capi.Handle(**cfg) as hd:
        fn = (lambda: hd.match_full(L, R)) if mode == 0 else (lambda: hd.match_foveated(L, R))
        for _ in range(3):
            out = fn()
        dev, wall = [], []
        for _ in range(reps):
            t0 = time.perf_counter()
            out = fn()
            wall.append((time.perf_counter() - t0) * 1e3)
            dev.append(hd.timing().total_ms)
        d = capi.query_dims(hd.cfg, h, w, mode)
    rec = {"config": name, "shape": [w, h], "mode": "foveated" if mode else "full", "levels": d.levels,
           "pixel_iterations": int(d.pixel_iterations), "device_ms_per_pair": float(np.median(dev)),
           "host_call_ms_per_pair": float(np.median(wall)), "launches": hd.timing().launches if False else None}
    if mode == 0:
        m = (slice(32, -32), slice(32, -32))
        rec["sanity_epe_h_px"] = float(np.abs(out[0] - dH)[m].mean())
    else:
        rec["fovea"] = [d.fov_width, d.fov_height]
        rec["stack_shape"] = list(out[0].shape)
    print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    run("configs[0] 640x480, 5-level pyramid", 640, 480, 0, levels=5)
    run("640x480, reference 14 levels", 640, 480, 0)
    run("configs[1] 4MP 2464x1632 full pyramid", 2464, 1632, 0)
    run("configs[2] 16MP 4928x3264 full pyramid", 4928, 3264, 0)
    run("configs[3] 16MP foveated (615x407 fovea)", 4928, 3264, 1)
