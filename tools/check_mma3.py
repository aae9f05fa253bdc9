#!/usr/bin/env python
"""Stage isolation: zero out parts of the decoder and compare FFMA vs tensor-core outputs."""
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
base_params = orc.init_params(spec, seed=3, density_gain=10.8, app_gain=3.0)
g = torch.Generator().manual_seed(5)
R = 1000
d = torch.cat([0.5 * (torch.rand(R, 2, generator=g) * 2 - 1), torch.ones(R, 1)], -1)
rays = torch.cat([0.02 * torch.randn(R, 3, generator=g), d / d.norm(dim=-1, keepdim=True)], -1)
S = orc.derive_step(spec)[1] // 2


def run(params, mode):
    os.environ["T2N_DECODER"] = mode
    m = build_model(spec, params, dev)
    with torch.no_grad():
        render_with_jitter(m, rays.to(dev), None, False, True, S)
    torch.cuda.synchronize()
    n = m.app_sample_count()[0]
    sc = m._last_scratch
    slots = sc["slots"][:n].cpu()
    rgb = sc["app_rgb"][:n].cpu()
    o = torch.argsort(slots)
    return rgb[o]


def variant(name, edit):
    p = {k: v.clone() for k, v in base_params.items()}
    edit(p)
    a, b = run(p, "ffma"), run(p, "mma")
    e = (a - b).abs()
    print(f"{name:28s} max {float(e.max()):.3e} mean {float(e.mean()):.3e}  ffma sample {a[0].tolist()} mma {b[0].tolist()}", flush=True)


W1, B1, W2, B2, W3 = ("renderModule.mlp.0.weight", "renderModule.mlp.0.bias", "renderModule.mlp.2.weight",
                      "renderModule.mlp.2.bias", "renderModule.mlp.4.weight")
variant("full", lambda p: None)
variant("W2=0 (L3+bias only)", lambda p: p[W2].zero_())
variant("W1=0 (L2,L3)", lambda p: p[W1].zero_())
variant("basis=0 (const L1 input)", lambda p: p["basis_mat.weight"].zero_())
variant("W1 trig cols=0", lambda p: p[W1][:, 27:].zero_())
variant("W1 ident cols=0", lambda p: p[W1][:, :27].zero_())
variant("b1=b2=0", lambda p: (p[B1].zero_(), p[B2].zero_()))
