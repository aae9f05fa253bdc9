#!/usr/bin/env python
"""Cycle-counter trace of the tensor-core appearance kernel (CTA 0) on the bench workload."""
import contextlib
import ctypes as C
import io
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["T2N_MMA_TRACE"] = "1"
import bench  # noqa: E402
from oracle import t2n_oracle as orc  # noqa: E402
from text2nerf_b200 import TensorVMSplit, _native as nat, ray_utils  # noqa: E402

dev = torch.device("cuda:0")
spec = bench.make_spec()
params = bench.make_params(spec)
S = orc.derive_step(spec)[1]
with contextlib.redirect_stdout(io.StringIO()):
    model = TensorVMSplit(spec.aabb_t().to(dev), bench.GRID, dev, density_n_comp=[16, 16, 16], appearance_n_comp=[48, 48, 48],
                          app_dim=27, near_far=bench.NEAR_FAR, shadingMode="MLP_Fea_noview", step_ratio=bench.STEP_RATIO,
                          fea_pe=6, view_pe=2)
model.load_state_dict({k: v.to(dev) for k, v in params.items()})
rays = ray_utils.camera_rays(bench.view_pose(0), bench.H, bench.W, [bench.FOCAL] * 2, device=dev)
for _ in range(2):
    with torch.no_grad():
        model(rays, is_train=False, white_bg=True, N_samples=S)
torch.cuda.synchronize()
buf = (C.c_longlong * 32)()
n = nat.load().t2n_debug_trace_read(buf)
v = list(buf)
nt = max(v[11], 1)
print("tiles of CTA0", v[11], "producer total cyc/tile", v[0] / nt)
print("per tile: S0 %.0f  S1 %.0f  S2 %.0f  S3 %.0f" % tuple(x / nt for x in v[1:5]))
print("per tile acc waits: D0 %.0f D1 %.0f D2 %.0f" % tuple(x / nt for x in v[5:8]))
print("per tile A-stage waits in S2 %.0f S1 %.0f S0 %.0f" % tuple(x / nt for x in v[8:11]))
print("issuer per tile: total %.0f  wait_B %.0f  wait_A %.0f  issue %.0f  prefetch(wait b_free) %.0f  chunks %d" %
      (v[16] / nt, v[17] / nt, v[18] / nt, v[19] / nt, v[20] / nt, v[21]))
