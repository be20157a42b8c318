import os, sys
sys.path.insert(0, ".")
import numpy as np
from ug_stereomatcher_b200 import capi, synth
L, R, _, _ = synth.make_pair(4928, 3264, 0)
base = None
for k1, k2 in [(400000, 700000), (200000, 700000), (400000, 200000), (200000, 200000), (600000, 700000), (400000, 1100000)]:
    os.environ["UGSM_K1_SMALL_PIXELS"], os.environ["UGSM_K2_SMALL_PIXELS"] = str(k1), str(k2)
    with capi.Handle() as h:
        for _ in range(3):
            out = h.match_foveated(L, R)
        fov = np.median([(h.match_foveated(L, R), h.timing().total_ms)[1] for _ in range(4)])
        for _ in range(2):
            h.match_full(L, R)
        full = np.median([(h.match_full(L, R), h.timing().total_ms)[1] for _ in range(3)])
    if base is None:
        base = [o.copy() for o in out[:3]]
    same = all(np.array_equal(a, b) for a, b in zip(out[:3], base))
    print("k1_small<%d k2_small<%d: foveated %.3f ms, full %.3f ms, same bits %s" % (k1, k2, fov, full, same), flush=True)
