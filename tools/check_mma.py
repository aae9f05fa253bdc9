#!/usr/bin/env python
"""A/B check of the tensor-core decoder against the exact FFMA decoder and the oracle."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import Case, build_model, render_with_jitter  # noqa: E402
from oracle import t2n_oracle as orc  # noqa: E402

dev = torch.device("cuda:0")


def run(spec, params, rays, jitter, train, S, mode):
    os.environ["T2N_DECODER"] = mode
    m = build_model(spec, params, dev)
    with torch.no_grad():
        out = render_with_jitter(m, rays.to(dev), jitter, train, True, S)
    torch.cuda.synchronize()
    return [t.cpu() for t in out], m.app_sample_count()


c = Case("t2n_noview_train")
for mode in ("ffma", "mma"):
    out, cnt = run(c.spec, c.params, c.rays, c.jitter, True, c.n_samples, mode)
    print(mode, "golden rgb max abs err", float((out[0] - c.out["rgb_map"]).abs().max()), "counts", cnt, flush=True)

spec = orc.FieldSpec(aabb=[[-8, -8, -8], [8, 8, 8]], grid=[64, 64, 64], near_far=[0.5, 8.0], step_ratio=1.0)
params = orc.init_params(spec, seed=3, density_gain=10.8, app_gain=3.0)
g = torch.Generator().manual_seed(5)
R = 3000
d = torch.cat([0.5 * (torch.rand(R, 2, generator=g) * 2 - 1), torch.ones(R, 1)], -1)
rays = torch.cat([0.02 * torch.randn(R, 3, generator=g), d / d.norm(dim=-1, keepdim=True)], -1)
S = orc.derive_step(spec)[1] // 2
ref = orc.render(spec, params, rays, S, False, True, None)
res = {}
for mode in ("ffma", "mma", "mma1", "mma3", "mma5"):
    if mode.startswith("mma") and len(mode) > 3:
        os.environ["T2N_MMA_TERMS"] = mode[3:]
    else:
        os.environ.pop("T2N_MMA_TERMS", None)
    out, cnt = run(spec, params, rays, None, False, S, mode[:3] if mode.startswith("mma") else mode)
    res[mode] = out
    err = (out[0] - ref[0]).abs()
    rel = (err / ref[0].abs().clamp_min(0.05)).max()
    print(mode, "64^3 rgb max abs", float(err.max()), "max rel(floor .05)", float(rel), "counts", cnt, flush=True)
print("mma vs ffma rgb max abs", float((res["mma"][0] - res["ffma"][0]).abs().max()))
