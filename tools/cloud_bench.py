#!/usr/bin/env python
"""Point-cloud step (K5) at BASELINE sizes on one B200: kernel time from CUDA events inside the library,
roofline fraction against MEASURED_PEAKS.json, and the pair + cloud path with the planes staying on the
device.  Prints JSON lines.  (The CPU implementations are timed beside it by the tests, which are allowed to
run the oracle: tests/test_gpu_cloud.py::test_16mp_full_cloud_vs_oracle and
tests/test_cloud_oracle.py::test_reference_binary_timing, both with `pytest -s`.)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np
import torch

from make_refpc_golden import cameras
from ug_stereomatcher_b200 import capi, synth


def peak_gbps():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        for k in ("hbm_gbs", "hbm_gbps", "hbm_GBps"):
            if k in p:
                return float(p[k]), "measured (MEASURED_PEAKS.json)"
        for v in p.values():
            if isinstance(v, dict):
                for k, x in v.items():
                    if "gb" in k.lower() and isinstance(x, (int, float)):
                        return float(x), "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    return 6553.9, "fallback"


def main():
    W, Hh = 4928, 3264
    rng = np.random.default_rng(1)
    P1, P2 = cameras(rng, W, Hh)
    cam = capi.CameraPair.of(P1, P2)
    L, R, _, _ = synth.make_pair(W, Hh, 0)
    peak, src = peak_gbps()
    dev = torch.device("cuda:0")
    tL, tR = torch.from_numpy(L).to(dev), torch.from_numpy(R).to(dev)
    outs = [torch.empty((Hh, W), dtype=torch.float32, device=dev) for _ in range(3)]
    with capi.Handle() as h:
        for _ in range(2):
            h.match_full_device(tL.data_ptr(), tR.data_ptr(), Hh, W, W * 3, *[o.data_ptr() for o in outs])
        pair_ms = h.timing().total_ms
        for stride, sampling in [(4, 1), (8, 1), (4, 5)]:
            n = capi.cloud_count(Hh, W, sampling)
            tP = torch.empty((n, stride), dtype=torch.float32, device=dev)
            ms = []
            for _ in range(8):
                h.cloud_from_match(cam, Hh, W, sampling=sampling, stride=stride, d_points=tP.data_ptr())
                ms.append(h.cloud_last_ms())
            k = float(np.median(ms[3:]))
            alg = n * (8 + 3 + 4 * stride)
            print(json.dumps({"what": "k5_cloud 16MP full, points stay on device", "sampling": sampling,
                              "stride_floats": stride, "points": n, "kernel_ms": k,
                              "algorithmic_bytes": alg, "achieved_GBps": alg / k / 1e6, "peak_GBps": peak,
                              "frac": alg / k / 1e6 / peak, "peak_source": src}), flush=True)
        # pair + cloud, device resident vs. pair + 3 planes to the host
        pin = torch.empty((Hh * W, 4), dtype=torch.float32).pin_memory()
        pts = pin.numpy()
        t = []
        for _ in range(5):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            h.match_full_device(tL.data_ptr(), tR.data_ptr(), Hh, W, W * 3, *[o.data_ptr() for o in outs])
            h.cloud_from_match(cam, Hh, W, out=pts)
            t.append((time.perf_counter() - t0) * 1e3)
        print(json.dumps({"what": "16MP pair (device-resident images) + cloud to pinned host memory (257 MB D2H)",
                          "pair_device_ms": pair_ms, "wall_ms": float(np.median(t[1:]))}), flush=True)


if __name__ == "__main__":
    main()
