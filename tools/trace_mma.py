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
for _ in range(2):
    with torch.no_grad():
        model(rays, is_train=False, white_bg=True, N_samples=S)
torch.cuda.synchronize()
N = 8192
buf = (C.c_longlong * N)()
n = nat.load().t2n_debug_trace_read_n(buf, N)
v = list(buf)
nt = max(v[11], 1)
print("tiles of CTA0", v[11], "producer total cyc/tile %.0f" % (v[0] / nt))
print("per tile: S2 %.0f  S1 %.0f  U(gather units) %.0f  Pre %.0f  Ray %.0f  Pro %.0f  S3 %.0f" % tuple(x / nt for x in v[1:8]))
print("per tile waits: accumulators %.0f  chunk-done (stage free) %.0f" % (v[8] / nt, v[9] / nt))
print("issuer per tile: total %.0f  wait_B %.0f  wait_A %.0f  issue %.0f  chunks %d" %
      (v[16] / nt, v[17] / nt, v[18] / nt, v[19] / nt, v[21]))

# ---- timeline of three iterations of CTA 0: ev[warp][iteration][chunk][kind]
NJ, CH = 3, 24
def E(w, jj, ci, k):
    return v[64 + ((w * NJ + jj) * CH + ci) * 4 + k]
t0 = min(x for x in v[64:64 + 18 * NJ * CH * 4] if x > 0)
kinds = ["S2"] * 4 + ["S1"] * 13 + ["U"] * 5
print("chunk timeline (cycles since first event).  producers: start = first warp starts the chunk, ready = last warp's data written,")
print("pub = last warp arrived;  issuer: B = weights landed, A = a_full seen, iss = MMAs issued;  loader: freeB = B stage free")
for jj in range(NJ):
    print("iteration", jj)
    for ci in range(22):
        st = [E(w, jj, ci, 0) for w in range(16) if E(w, jj, ci, 0) > 0]
        st3 = [E(w, jj, ci, 3) for w in range(16) if E(w, jj, ci, 3) > 0]
        rd = [E(w, jj, ci, 1) for w in range(16) if E(w, jj, ci, 1) > 0]
        pb = [E(w, jj, ci, 2) for w in range(16) if E(w, jj, ci, 2) > 0]
        if not pb:
            continue
        print("  c%02d start %6d..%6d  k3 %6d..%6d  k1 %6d..%6d  pub %6d..%6d | issuer B %6d A %6d iss %6d | loader freeB %6d" % (
            ci, min(st) - t0 if st else -1, max(st) - t0 if st else -1, min(st3) - t0 if st3 else -1, max(st3) - t0 if st3 else -1,
            min(rd) - t0 if rd else -1, max(rd) - t0 if rd else -1, min(pb) - t0, max(pb) - t0,
            E(16, jj, ci, 0) - t0, E(16, jj, ci, 1) - t0, E(16, jj, ci, 2) - t0, E(17, jj, ci, 0) - t0))

# per-warp lateness: mean over chunks of (publish time of the warp - earliest publish of the chunk)
late = [0.0] * 16
cnt = 0
for jj in range(NJ):
    for ci in range(22):
        pb = [E(w, jj, ci, 2) for w in range(16)]
        if min(pb) <= 0:
            continue
        m0 = min(pb)
        for w in range(16):
            late[w] += pb[w] - m0
        cnt += 1
print("mean publish lateness per producer warp (cycles):", " ".join("%d:%.0f" % (w, late[w] / max(cnt, 1)) for w in range(16)))
