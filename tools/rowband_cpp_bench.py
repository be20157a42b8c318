"""Row-band split of ONE 16 MP pair over the GPUs of a box through the C++ driver ugsm_rowband_match (single process,
one host thread per GPU, NVLink peer copies).  python tools/rowband_cpp_bench.py [n_gpus ...]  -> one JSON line per N,
each checked bit for bit against the single-GPU result."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ug_stereomatcher_b200 import capi, synth

COLS, ROWS = 4928, 3264
L, R, _, _ = synth.make_pair(COLS, ROWS, 0)
ngpu = torch.cuda.device_count()
worlds = [int(a) for a in sys.argv[1:]] or [n for n in (1, 2, 4, 8) if n <= ngpu]
with capi.Handle(device=0) as h:
    want = h.match_full(L, R)
    single = min((h.match_full(L, R), h.timing().total_ms)[1] for _ in range(3))
for n in worlds:
    devs = list(range(n))
    dL = [torch.from_numpy(L).to("cuda:%d" % d) for d in devs]
    dR = [torch.from_numpy(R).to("cuda:%d" % d) for d in devs]
    out = [torch.empty((ROWS, COLS), dtype=torch.float32, device="cuda:0") for _ in range(3)]
    hs = [capi.Handle(device=d) for d in devs]
    ms = []
    for i in range(8):
        ms.append(capi.rowband_match(hs, devs, [t.data_ptr() for t in dL], [t.data_ptr() for t in dR], ROWS, COLS,
                                     COLS * 3, [o.data_ptr() for o in out]))
    torch.cuda.synchronize()
    same = all((o.cpu().numpy() == w).all() for o, w in zip(out, want))
    print(json.dumps({"mode": "rowband (C++ driver, peer copies)", "n_gpus": n, "ms_per_pair": float(np.median(ms[3:])),
                      "ms_all": [round(x, 3) for x in ms], "single_gpu_ms": single, "bit_exact_vs_single_gpu": bool(same)}), flush=True)
    for hd in hs:
        hd.close()
