#!/usr/bin/env python
"""Cycle-counter trace of app_backward_kernel (CTA 0) on one 4096-ray training batch of the bench workload."""
import contextlib
import ctypes as C
import io
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["T2N_BWD_TRACE"] = "1"
import bench  # noqa: E402
from text2nerf_b200 import TensorVMSplit, _native as nat, ray_utils  # noqa: E402

dev = torch.device("cuda:0")
params = bench.make_params()
with contextlib.redirect_stdout(io.StringIO()):
    model = TensorVMSplit(torch.tensor(bench.AABB, dtype=torch.float32, device=dev), bench.GRID, dev, density_n_comp=[16, 16, 16], appearance_n_comp=[48, 48, 48],
                          app_dim=27, near_far=bench.NEAR_FAR, shadingMode="MLP_Fea_noview", step_ratio=bench.STEP_RATIO,
                          fea_pe=6, view_pe=2)
model.load_state_dict({k: v.to(dev) for k, v in params.items()})
S = model.nSamples
rays = ray_utils.camera_rays(bench.view_pose(0), bench.H, bench.W, [bench.FOCAL] * 2, device=dev)
g = torch.Generator().manual_seed(0)
for _ in range(2):
    idx = torch.randint(0, rays.shape[0], (4096,), generator=g).to(dev)
    out = model(rays[idx].contiguous(), is_train=True, white_bg=True, N_samples=S)
    bench.composed_loss(*out, torch.rand(4096, 3, generator=g).to(dev), (2 + 4 * torch.rand(4096, generator=g)).to(dev)).backward()
torch.cuda.synchronize()
buf = (C.c_longlong * 32)()
nat.load().t2n_debug_trace_read(buf)
v = list(buf)
nt = max(v[12], 1)
if os.environ.get("T2N_BWD_FFMA"):
    names = ["gather+basis", "decoder fwd recompute", "L3 fwd/bwd,dW3,dz2", "dW2+red", "dh1->dz1", "L1bwd: load+columns", "dW1+red",
             "dA+PE bwd", "basis bwd", "scatter", "-", "loop overhead"]
    nt = max(v[12], 1)
    print("FFMA kernel, tiles of CTA0:", v[12], " total cycles/tile:", sum(v[:12]) / nt)
    tot = sum(v[:12])
else:
    names = ["P1 dz3,dz2 -> TMEM + images", "wait dh1 MMA", "P2 dz1 -> TMEM + image", "P3 dA ring: PE backward + column images",
             "P4 dfeat", "P5 dprod -> global (gather + scatter: app_scatter_kernel)"]
    nt = max(v[8], 1)
    print("tensor-core backward-data kernel, tiles (128 samples) of CTA0:", v[8], " total cycles/tile:", sum(v[:6]) / nt)
    tot = sum(v[:6])
for n, c in zip(names, v[:len(names)]):
    print(f"  {n:44s} {c / nt:10.0f} cyc/tile  {100 * c / max(tot, 1):5.1f}%")
