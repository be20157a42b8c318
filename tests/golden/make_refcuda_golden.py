#!/usr/bin/env python
"""Generate tests/golden/refcuda_1232x816.npz: outputs of the UNMODIFIED reference CUDA matcher
(oracle/_ref/refcuda, built by oracle/refbuild/Makefile from /root/reference through compat shims)
run on a B200, on the synthetic pair synth.make_pair(1232, 816, 0), reference defaults (14 levels).

Run on the GPU box:   python tests/golden/make_refcuda_golden.py gpurun_out [W H]
(W H = 2464 1632 repeats the run at BASELINE configs[1]'s size: statistics and timings only,
tests/golden/refcuda_2464x1632_stats.json).  It stores two 320x240 crops (one interior, one touching the top-left border) of H, V, C plus the
full-frame agreement statistics against the oracle and against this repo's CUDA path, so the CPU
test (tests/test_refcuda_golden.py) can pin the oracle without a GPU.
"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

W, H = 1232, 816
CROPS = {"interior": (300, 450, 240, 320), "corner": (0, 0, 240, 320)}  # y0, x0, h, w


def main(outdir, W=W, H=H):
    import ugsm_oracle as oracle
    from ug_stereomatcher_b200 import capi, synth

    L, R, dH, dV = synth.make_pair(W, H, 0)
    tmp = "/tmp/refcuda_%d" % os.getpid()
    os.makedirs(tmp, exist_ok=True)
    L.tofile(tmp + "/l.rgb")
    R.tofile(tmp + "/r.rgb")
    exe = os.path.join(ROOT, "oracle", "_ref", "refcuda")
    t0 = time.time()
    p = subprocess.run([exe, str(H), str(W), tmp + "/l.rgb", tmp + "/r.rgb", tmp + "/out.f32"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=1500)
    t_ref = time.time() - t0
    log = p.stdout.decode(errors="ignore")
    if p.returncode != 0:
        print(log[-3000:])
        raise SystemExit("reference CUDA build failed to run: rc=%d" % p.returncode)
    out = np.fromfile(tmp + "/out.f32", np.float32).reshape(3, H, W)
    t0 = time.time()
    oH, oV, oC = oracle.match_full(L, R)
    t_or = time.time() - t0
    with capi.Handle() as h:
        gH, gV, gC = h.match_full(L, R)
        t_gpu = h.timing().total_ms

    def stats(a, b, name):
        d = np.abs(a.astype(np.float64) - b)
        inner = d[16:-16, 16:-16]
        return {name + "_pct_within_0.05": float((d <= 0.05).mean() * 100),
                name + "_pct_within_0.05_interior16": float((inner <= 0.05).mean() * 100),
                name + "_pct_within_1e-3": float((d <= 1e-3).mean() * 100),
                name + "_pct_bit_equal": float((a == b).mean() * 100),
                name + "_max_abs": float(np.nanmax(d)), name + "_median_abs": float(np.median(d))}

    st = {"shape": [H, W], "reference_seconds": t_ref, "oracle_seconds": t_or, "gpu_ms": t_gpu,
          "nonfinite_ref": int((~np.isfinite(out)).sum())}
    for k, (ref, orc, gpu) in {"H": (out[0], oH, gH), "V": (out[1], oV, gV), "C": (out[2], oC, gC)}.items():
        st.update(stats(ref, orc, "ref_vs_oracle_" + k))
        st.update(stats(ref, gpu, "ref_vs_gpu_" + k))
        st["gpu_vs_oracle_bit_equal_" + k] = bool((gpu == orc).all())
    m = (slice(32, -32), slice(32, -32))
    st["epe_ref_H"] = float(np.abs(out[0] - dH)[m].mean())
    st["epe_oracle_H"] = float(np.abs(oH - dH)[m].mean())
    arrays = {}
    for name, (y0, x0, hh, ww) in CROPS.items():
        arrays[name] = out[:, y0:y0 + hh, x0:x0 + ww].copy()
    os.makedirs(outdir, exist_ok=True)
    tag = "refcuda_%dx%d" % (W, H)
    if (W, H) == (1232, 816):
        np.savez_compressed(os.path.join(outdir, tag + ".npz"), **arrays)
    json.dump(st, open(os.path.join(outdir, tag + "_stats.json"), "w"), indent=1)
    open(os.path.join(outdir, tag + "_log_tail.txt"), "w").write(log[-4000:])
    print(json.dumps(st, indent=1))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out", *(int(a) for a in sys.argv[2:4]))
