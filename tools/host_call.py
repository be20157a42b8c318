#!/usr/bin/env python
"""Wall time of ugsm_match_full with PAGEABLE host buffers (what the MatchGPULib facade hands out) at 16 MP,
for a few UGSM_HOST_THREADS settings (0 = plain cudaMemcpyAsync from/to the caller's memory)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from ug_stereomatcher_b200 import capi, synth

L, R, _, _ = synth.make_pair(4928, 3264, 0)
ref = None
for threads in (0, 4, 8):
    os.environ["UGSM_HOST_THREADS"] = str(threads)
    with capi.Handle() as h:
        fresh, reuse = [], []
        for i in range(5):
            t0 = time.perf_counter()
            out = h.match_full(L, R)  # fresh output arrays every call, like the facade's malloc
            fresh.append((time.perf_counter() - t0) * 1e3)
        for i in range(4):
            t0 = time.perf_counter()
            h.match_full(L, R, out=out)  # caller reuses its planes
            reuse.append((time.perf_counter() - t0) * 1e3)
        dev = h.timing().total_ms
    if ref is None:
        ref = [o.copy() for o in out]
    same = all(np.array_equal(a, b) for a, b in zip(out, ref))
    print("threads %2d: fresh outputs %.1f ms, reused outputs %.1f ms (device %.2f ms), identical to plain path: %s" % (
        threads, float(np.median(fresh[1:])), float(np.median(reuse[1:])), dev, same), flush=True)
