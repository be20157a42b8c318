#!/usr/bin/env python
"""Two 16 MP pairs through the host C-ABI (for an ncu launch list of the second one:
UGSM_USE_GRAPH=0 ncu --launch-skip 479 -c 479 --metrics gpu__time_duration.sum --csv python tools/one_pair.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ug_stereomatcher_b200 import capi, synth

w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4928, 3264)
L, R, _, _ = synth.make_pair(w, h, 0)
with capi.Handle() as hd:
    for _ in range(2):
        hd.match_full(L, R)
    print("device ms", hd.timing().total_ms, "launches", hd.timing().launches)
