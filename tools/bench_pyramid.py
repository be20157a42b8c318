#!/usr/bin/env python
"""K0 (pyramid / left-energy) launch times per level size; rows per warp of the marching kernels via
UGSM_K0_ROWS_DENSE / UGSM_K0_ROWS_DEC."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ug_stereomatcher_b200 import capi

h = capi.Handle()
print("rows/warp dense=%s dec=%s" % (os.environ.get("UGSM_K0_ROWS_DENSE", "default"), os.environ.get("UGSM_K0_ROWS_DEC", "default")))
for (w, hh) in [(4928, 3264), (3484, 2307), (2463, 1631), (1741, 1153), (1231, 815)]:
    print("%5dx%-5d energy %.4f ms  dec2 %.4f ms  sqrt2 %.4f ms" % (
        w, hh, h.bench_pyramid(w, hh, 0), h.bench_pyramid(w, hh, 1), h.bench_pyramid(w, hh, 2)))
