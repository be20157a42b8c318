import os, sys, time
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as np
from ug_stereomatcher_b200 import capi, synth
L, R, _, _ = synth.make_pair(4928, 3264, 0)
res = {}
for g in (0, 1):
    h = capi.Handle(use_graph=g)
    for i in range(4):
        t0 = time.perf_counter(); out = h.match_full(L, R); t1 = time.perf_counter()
        t = h.timing()
    res[g] = [o.copy() for o in out]
    print("graph=%d: host-call %.1f ms, device %.2f ms (pyramid %.2f, levels %.2f) launches %d" % (g, (t1 - t0) * 1e3, t.total_ms, t.pyramid_ms, t.levels_ms, t.launches))
    print("   level ms", [round(x, 3) for x in t.level_ms])
    print("   k1 ms l0..3", [round(x, 3) for x in t.k1_ms[:4]], "k2", [round(x, 3) for x in t.k2_ms[:4]])
    h.close()
print("graph == no-graph bits:", all((a == b).all() for a, b in zip(res[0], res[1])))
