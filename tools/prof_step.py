#!/usr/bin/env python
"""Small driver for ncu captures: the bench workload (lego800) with a configurable number of
forward views and training batches, no timing.  Used only under the profiler."""
import argparse
import contextlib
import io
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--views", type=int, default=1)
    ap.add_argument("--rays", type=int, default=bench.H * bench.W)
    ap.add_argument("--train-batches", type=int, default=1)
    ap.add_argument("--batch", type=int, default=bench.TRAIN_BATCH)
    a = ap.parse_args()
    from text2nerf_b200 import TensorVMSplit, ray_utils
    dev = torch.device("cuda:0")
    params = bench.make_params()
    with contextlib.redirect_stdout(io.StringIO()):
        model = TensorVMSplit(torch.tensor(bench.AABB, dtype=torch.float32, device=dev), bench.GRID, dev, density_n_comp=[16, 16, 16],
                              appearance_n_comp=[48, 48, 48], app_dim=27, near_far=bench.NEAR_FAR,
                              shadingMode="MLP_Fea_noview", step_ratio=bench.STEP_RATIO, fea_pe=6, view_pe=2)
    model.load_state_dict({k: v.to(dev) for k, v in params.items()})
    S = model.nSamples
    rays = ray_utils.camera_rays(bench.view_pose(0), bench.H, bench.W, [bench.FOCAL] * 2, device=dev)[:a.rays].contiguous()
    for _ in range(a.views):
        with torch.no_grad():
            model(rays, is_train=False, white_bg=True, N_samples=S)
    g = torch.Generator().manual_seed(0)
    for _ in range(a.train_batches):
        idx = torch.randint(0, rays.shape[0], (a.batch,), generator=g).to(dev)
        loss = model.data_loss(rays[idx].contiguous(), torch.rand(a.batch, 3, generator=g).to(dev),
                               (2 + 4 * torch.rand(a.batch, generator=g)).to(dev), white_bg=True, N_samples=S)
        loss.backward()
    torch.cuda.synchronize()
    print("done", model.app_sample_count())


if __name__ == "__main__":
    main()
