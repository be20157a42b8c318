"""One K2 micro-benchmark run at a given size (for ncu captures): python tools/k2_one.py [w h passes reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ug_stereomatcher_b200 import capi
a = sys.argv[1:5] + ["4928", "3264", "10", "3"][len(sys.argv) - 1:]
w, hh, s, reps = (int(x) for x in a)
h = capi.Handle()
print("K2 %dx%d S=%d: %.4f ms" % (w, hh, s, h.bench_smooth(w, hh, s, reps)))
