import sys
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from ug_stereomatcher_b200 import capi
h = capi.Handle()
print("null launch us", h.bench_null(200) * 1e3)
for (w, hh) in [(615, 407), (216, 142), (53, 34)]:
    print(w, hh, "K1 us %.1f K2S5 us %.1f" % (h.bench_correlate(w, hh, 20) * 1e3, h.bench_smooth(w, hh, 5, 20) * 1e3))
