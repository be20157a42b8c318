#!/usr/bin/env python
"""Minimal workload for an `ncu --set full` capture of the HBM-bound kernels at 16 MP: K0 (left energy, blur + x2
decimation) and K5 (point cloud):
  ncu --set full --clock-control none -k "regex:k0_march5|k5_cloud" -c 8 -o out python tools/prof_hbm_kernels.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np
import torch

from make_refpc_golden import cameras
from ug_stereomatcher_b200 import capi

W, H = 4928, 3264
h = capi.Handle()
print("energy", h.bench_pyramid(W, H, 0, 2), "dec2", h.bench_pyramid(W, H, 1, 2))
rng = np.random.default_rng(0)
P1, P2 = cameras(rng, W, H)
dev = torch.device("cuda:0")
dX = torch.randn((H, W), device=dev) * 5
dY = torch.randn((H, W), device=dev)
img = torch.randint(0, 256, (H, W, 3), dtype=torch.uint8, device=dev)
pts = torch.empty((H * W, 4), dtype=torch.float32, device=dev)
torch.cuda.synchronize()
for _ in range(3):
    h.cloud_full_device(capi.CameraPair.of(P1, P2), dX.data_ptr(), dY.data_ptr(), W, H, W, img.data_ptr(), W * 3, True,
                        1, pts.data_ptr())
print("cloud", h.cloud_last_ms())
