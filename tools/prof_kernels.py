import sys
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from ug_stereomatcher_b200 import capi
h = capi.Handle()
w, hh = 4928, 3264
print("K1", h.bench_correlate(w, hh, 2), "K2 S10", h.bench_smooth(w, hh, 10, 2), "K2 S5", h.bench_smooth(w, hh, 5, 2))
print("small K1", h.bench_correlate(216, 142, 2), "small K2", h.bench_smooth(216, 142, 5, 2))
