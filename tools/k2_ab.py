"""A/B of the K2 kernels on a B200: level sizes of a 16 MP pair, tiled k2_smooth_box vs k2_stream."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ug_stereomatcher_b200 import capi

SIZES = [(4928, 3264, 10), (3484, 2307, 10), (2463, 1631, 5), (1741, 1153, 5), (1231, 815, 5), (870, 576, 5), (615, 407, 5)]
extra = dict(kv.split("=") for kv in sys.argv[1:])
for name, env in [("tiled", dict(UGSM_K2_STREAM=0)), ("stream", dict(UGSM_K2_STREAM=1, UGSM_K2S_MIN_PIXELS=0, UGSM_K2S_MAX_PASSES=10))]:
    os.environ.update({k: str(v) for k, v in env.items()})
    os.environ.update(extra)
    os.environ["UGSM_K2_SMALL_PIXELS"] = "0"
    h = capi.Handle()
    row = ["%-8s" % name]
    for (w, hh, s) in SIZES:
        ms = h.bench_smooth(w, hh, s, 10)
        row.append("%dx%d S%d %.4f ms (%.0f GB/s)" % (w, hh, s, ms, 24.0 * w * hh / ms / 1e6))
    print(" | ".join(row), flush=True)
    h.close()
