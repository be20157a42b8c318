import os, sys, time
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as np
from ug_stereomatcher_b200 import capi, synth
h = capi.Handle(use_graph=2)  # level 0 outside the graph: per-launch K1 / K2 times
for (w, hh) in [(4928, 3264), (2463, 1631), (870, 576), (615, 407), (216, 142), (53, 34)]:
    print("%5dx%-5d K1 %.4f ms  K2(S5) %.4f ms  K2(S10) %.4f ms" % (w, hh, h.bench_correlate(w, hh, 5), h.bench_smooth(w, hh, 5, 5), h.bench_smooth(w, hh, 10, 5)))
if os.environ.get("PAIR", "1") == "1":
    L, R, _, _ = synth.make_pair(4928, 3264, 0)
    for i in range(3):
        t0 = time.perf_counter(); h.match_full(L, R); t1 = time.perf_counter()
        t = h.timing()
        print("pair %d: host-call %.1f ms, device %.2f ms (pyramid %.2f, levels %.2f) launches %d" % (i, (t1 - t0) * 1e3, t.total_ms, t.pyramid_ms, t.levels_ms, t.launches))
    print("level ms", [round(x, 3) for x in t.level_ms])
    print("k1 ms l0..3", [round(x, 3) for x in t.k1_ms[:4]], "k2", [round(x, 3) for x in t.k2_ms[:4]])
