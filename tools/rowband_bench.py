#!/usr/bin/env python
"""Single-pair latency with the row-band split across GPUs, one process per GPU:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/rowband_bench.py
Checks on rank 0 that the assembled result is bit-identical to the single-GPU path, then times
`--steps` pairs (CUDA events, max over ranks).  Prints one JSON line."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from ug_stereomatcher_b200 import capi, dp, rowband, synth

ap = argparse.ArgumentParser()
ap.add_argument("--cols", type=int, default=4928)
ap.add_argument("--rows", type=int, default=3264)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--split-levels", type=int, default=3)
a = ap.parse_args()

rank, world, local = dp.env_rank()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dp.init("nccl", dev)
L, R, _, _ = synth.make_pair(a.cols, a.rows, 0)
dL, dR = torch.from_numpy(L).to(dev), torch.from_numpy(R).to(dev)
out = [torch.empty((a.rows, a.cols), dtype=torch.float32, device=dev) for _ in range(3)]
h = capi.Handle(device=local)


def step():
    if world > 1:
        rowband.run_distributed(h, dL.data_ptr(), dR.data_ptr(), a.rows, a.cols, a.cols * 3, out, a.split_levels)
    else:
        rowband.run_emulated([h], dL.data_ptr(), dR.data_ptr(), a.rows, a.cols, a.cols * 3, out, a.split_levels)


for _ in range(3):
    step()
torch.cuda.synchronize()
ok = None
if rank == 0:
    with capi.Handle(device=local) as ref:
        want = ref.match_full(L, R)
        single_ms = ref.timing().total_ms
        for _ in range(3):
            ref.match_full(L, R)
            single_ms = min(single_ms, ref.timing().total_ms)
    ok = all(bool((o.cpu().numpy() == w).all()) for o, w in zip(out, want))
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    step()
e1.record()
torch.cuda.synchronize()
ms = dp.max_over_ranks(e0.elapsed_time(e1) / a.steps, dev)
if rank == 0:
    print(json.dumps({"what": "single-pair latency, row-band split of levels < %d" % a.split_levels, "n_gpus": world,
                      "shape": [a.cols, a.rows], "ms_per_pair_rowband": ms, "ms_per_pair_single_gpu_device": single_ms,
                      "speedup_vs_single_gpu": single_ms / ms, "bit_identical_to_single_gpu": ok,
                      "halo_rows": rowband.HALO, "exchanges_per_pair": sum([2, 4, 6][: a.split_levels])}))
if world > 1:
    dist.destroy_process_group()
