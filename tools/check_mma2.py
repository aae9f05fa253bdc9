#!/usr/bin/env python
"""Per-sample comparison of the decoder output (app_rgb) between the FFMA and tensor-core kernels."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import build_model, render_with_jitter  # noqa: E402
from oracle import t2n_oracle as orc  # noqa: E402

os.environ["T2N_KEEP_SCRATCH"] = "1"
dev = torch.device("cuda:0")
spec = orc.FieldSpec(aabb=[[-8, -8, -8], [8, 8, 8]], grid=[64, 64, 64], near_far=[0.5, 8.0], step_ratio=1.0)
params = orc.init_params(spec, seed=3, density_gain=10.8, app_gain=3.0)
g = torch.Generator().manual_seed(5)
R = 3000
d = torch.cat([0.5 * (torch.rand(R, 2, generator=g) * 2 - 1), torch.ones(R, 1)], -1)
rays = torch.cat([0.02 * torch.randn(R, 3, generator=g), d / d.norm(dim=-1, keepdim=True)], -1)
S = orc.derive_step(spec)[1] // 2
res = {}
for mode in ("ffma", "mma"):
    os.environ["T2N_DECODER"] = mode
    m = build_model(spec, params, dev)
    with torch.no_grad():
        out = render_with_jitter(m, rays.to(dev), None, False, True, S)
    torch.cuda.synchronize()
    n = m.app_sample_count()[0]
    sc = m._last_scratch
    slots = sc["slots"][:n].cpu()
    rgb = sc["app_rgb"][:n].cpu()
    order = torch.argsort(slots)
    res[mode] = (slots[order], rgb[order], order)
sa, ra, _ = res["ffma"]
sb, rb, ob = res["mma"]
assert torch.equal(sa, sb)
err = (ra - rb).abs().max(-1).values
print("n", len(err), "max", float(err.max()), "mean", float(err.mean()), "frac>1e-4", float((err > 1e-4).float().mean()),
      "frac>1e-3", float((err > 1e-3).float().mean()))
# position of bad points inside their 128-tile in the mma run's list order
pos = ob  # list index e of each sorted sample in the mma run
bad = err > 1e-3
print("bad count", int(bad.sum()))
if bad.any():
    e = pos[bad]
    print("bad list idx %128:", sorted((e % 128).tolist())[:40])
    print("bad tiles:", sorted(set((e // 128).tolist()))[:40], "of", (len(err) + 127) // 128)
q = torch.tensor([0.5, 0.9, 0.99, 0.999])
print("quantiles", torch.quantile(err, q).tolist())
worst = torch.argsort(err, descending=True)[:5]
for w in worst:
    print("worst", int(sa[w]), ra[w].tolist(), rb[w].tolist())
