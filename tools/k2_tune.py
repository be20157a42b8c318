"""k2_stream tuning on a B200: launch times at the level sizes of a 16 MP pair for the library UGSM_LIB points at
(build variants: make OUT=... EXTRA_NVFLAGS=-DK2S_MINB_SMALL=3) and the environment overrides given as K=V args."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.update(dict(kv.split("=") for kv in sys.argv[1:]))
for k, v in dict(UGSM_K2_STREAM="1", UGSM_K2S_MIN_PIXELS="0", UGSM_K2S_MAX_PASSES="10", UGSM_K2_SMALL_PIXELS="0").items():
    os.environ.setdefault(k, v)
from ug_stereomatcher_b200 import capi
SIZES = [(4928, 3264, 10), (3484, 2307, 10), (4928, 3264, 5), (2463, 1631, 5), (1741, 1153, 5), (1231, 815, 5), (870, 576, 5)]
if os.environ.get("K2_SIZES"):  # e.g. K2_SIZES=615x407x5,434x287x5
    SIZES = [tuple(int(v) for v in t.split("x")) for t in os.environ["K2_SIZES"].split(",")]
h = capi.Handle()
print("%-50s" % (os.path.basename(os.environ.get("UGSM_LIB", "default")) + " " + " ".join(sys.argv[1:])),
      " | ".join("%dx%d S%d %.4f" % (w, hh, s, h.bench_smooth(w, hh, s, 10)) for (w, hh, s) in SIZES), flush=True)
