#!/usr/bin/env python
"""Times the TV regularisers of the training step (tensoRF.py:193-203 with utils.TVLoss) as tensor ops on the
300^3 field, forward + backward, to size SURVEY 8f rank 2."""
import contextlib, io, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from text2nerf_b200 import TensorVMSplit

dev = torch.device("cuda:0")
with contextlib.redirect_stdout(io.StringIO()):
    model = TensorVMSplit(torch.tensor(bench.AABB, dtype=torch.float32, device=dev), bench.GRID, dev, density_n_comp=[16, 16, 16], appearance_n_comp=[48, 48, 48],
                          app_dim=27, near_far=bench.NEAR_FAR, shadingMode="MLP_Fea_noview", step_ratio=bench.STEP_RATIO)
flat = model.enable_flat_grads(True)
tvreg = bench.TVLoss().forward        # a plain callable: the tensor-op route (the TVLoss module itself takes the fused kernels)


def step():
    flat.zero_()
    loss = model.TV_loss_density(tvreg) * 0.1 + model.TV_loss_app(tvreg) * 0.01
    loss.backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    step()
e1.record()
torch.cuda.synchronize()
print("TV density + app, forward + backward, tensor ops: %.3f ms per step" % (e0.elapsed_time(e1) / 10))
opt = torch.optim.Adam(model.get_optparam_groups(0.02, 0.001), betas=(0.9, 0.99))
step(); opt.step()
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    opt.step()
e1.record()
torch.cuda.synchronize()
print("torch.optim.Adam step over the 19 tensors: %.3f ms" % (e0.elapsed_time(e1) / 10))
