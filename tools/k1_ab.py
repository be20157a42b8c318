"""A/B of the K1 kernels on a B200: level sizes of a 16 MP pair, tiled k1_correlate vs k1_stream (GENERIC / FAST)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ug_stereomatcher_b200 import capi

SIZES = [(4928, 3264), (3484, 2307), (2463, 1631), (1741, 1153), (1231, 815), (870, 576), (615, 407), (434, 287)]
VARIANTS = [("tiled", dict(UGSM_K1_STREAM=0)), ("stream generic", dict(UGSM_K1_STREAM=1, UGSM_K1_FAST=0)),
            ("stream fast", dict(UGSM_K1_STREAM=1, UGSM_K1_FAST=1))]
if os.environ.get("K1_SIZES"):  # e.g. K1_SIZES=615x407,434x287
    SIZES = [tuple(int(v) for v in t.split("x")) for t in os.environ["K1_SIZES"].split(",")]
extra = dict(kv.split("=") for kv in sys.argv[1:])
for name, env in VARIANTS:
    os.environ.update({k: str(v) for k, v in env.items()})
    os.environ.update(extra)
    os.environ["UGSM_K1_SMALL_PIXELS"] = "0"
    h = capi.Handle()
    row = ["%-15s" % name]
    for (w, hh) in SIZES:
        ms = h.bench_correlate(w, hh, 10)
        row.append("%dx%d %.4f ms (%.0f GB/s)" % (w, hh, ms, 48.0 * w * hh / ms / 1e6))
    print(" | ".join(row), flush=True)
    h.close()
