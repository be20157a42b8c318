"""One K1 micro-benchmark run at a given size (for ncu captures): python tools/k1_one.py [w h reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ug_stereomatcher_b200 import capi
w, hh, reps = (int(a) for a in (sys.argv[1:4] + ["4928", "3264", "3"][len(sys.argv) - 1:]))
h = capi.Handle()
print("K1 %dx%d: %.4f ms" % (w, hh, h.bench_correlate(w, hh, reps)))
