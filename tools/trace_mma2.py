#!/usr/bin/env python
"""Wait-time breakdown of the role-specialised appearance kernel (CTA 0) on the bench workload (T2N_V2_TRACE)."""
import contextlib
import ctypes as C
import io
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["T2N_V2_TRACE"] = "1"
import bench  # noqa: E402
from text2nerf_b200 import TensorVMSplit, _native as nat, ray_utils  # noqa: E402

dev = torch.device("cuda:0")
params = bench.make_params()
with contextlib.redirect_stdout(io.StringIO()):
    model = TensorVMSplit(torch.tensor(bench.AABB, dtype=torch.float32, device=dev), bench.GRID, dev, density_n_comp=[16, 16, 16],
                          appearance_n_comp=[48, 48, 48], app_dim=27, near_far=bench.NEAR_FAR, shadingMode="MLP_Fea_noview",
                          step_ratio=bench.STEP_RATIO, fea_pe=6, view_pe=2)
model.load_state_dict({k: v.to(dev) for k, v in params.items()})
S = model.nSamples
rays = ray_utils.camera_rays(bench.view_pose(0), bench.H, bench.W, [bench.FOCAL] * 2, device=dev)
for _ in range(2):
    with torch.no_grad():
        model(rays, is_train=False, white_bg=True, N_samples=S)
torch.cuda.synchronize()
buf = (C.c_longlong * 64)()
nat.load().t2n_debug_trace_read_n(buf, 64)
v = list(buf)
nt = max(v[4], 1)
print("tiles of CTA 0:", v[4])
print("decoder warp 0 : %6.0f cyc/tile | wait D1 %6.0f  wait D0 %6.0f  wait A stage %6.0f" % tuple(x / nt for x in v[0:4]))
print("decoder issuer : %6.0f cyc/tile | wait A chunk %6.0f  wait weights %6.0f" % tuple(x / nt for x in v[8:11]))
print("gather warp 8  : %6.0f cyc/tile | wait stage %6.0f" % tuple(x / nt for x in v[16:18]))
print("basis issuer   : %6.0f cyc/tile | wait A chunk %6.0f  wait weights %6.0f  wait D0 free %6.0f" % tuple(x / nt for x in v[24:28]))
print("decoder warp 0 phases per tile: S2 %6.0f  chunk 0 %6.0f  seeds %6.0f  PE chunks %6.0f  S3 %6.0f" % tuple(x / nt for x in v[32:37]))
print("decoder warp 0: tcgen05.wait::st %6.0f cyc/tile" % (v[5] / nt))
