"""Diagnostic: does the tensor-core backward mishandle the last (partial) 128-sample tile?  Compares TC vs FFMA backward."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["T2N_KEEP_SCRATCH"] = "1"
import torch
from oracle import t2n_oracle as orc
from helpers import build_model, render_with_jitter, scaled_err

dev = torch.device("cuda:0")
spec = orc.FieldSpec(aabb=[[-1.5, -1.5, 2.5], [1.5, 1.5, 5.5]], grid=[300, 300, 300], near_far=[2.0, 6.0], step_ratio=0.5)
params = orc.init_params(spec, seed=0, density_gain=10.8, app_gain=1.0)
S = orc.derive_step(spec)[1]
model = build_model(spec, params, dev)
for R in [int(x) for x in os.environ.get("RS", "64,128,256,512,1024,3000").split(",")]:
    g = torch.Generator().manual_seed(9)
    px = torch.rand(R, 2, generator=g) * 800.0
    d = torch.cat([(px - 400.0) / 1111.1, torch.ones(R, 1)], -1)
    rays = torch.cat([torch.zeros(R, 3), d / d.norm(dim=-1, keepdim=True)], -1).contiguous().to(dev)
    jitter = torch.rand(R, 1, generator=g)
    rgb_gt = torch.rand(R, 3, generator=g)
    depth_gt = 2.0 + 4.0 * torch.rand(R, generator=g)
    res = {}
    for mode in ("mma", "ffma"):
        if mode == "ffma":
            os.environ["T2N_BWD_FFMA"] = "1"
        else:
            os.environ.pop("T2N_BWD_FFMA", None)
        for rep in range(2):
            model.zero_grad()
            out = render_with_jitter(model, rays, jitter, True, True, S)
            loss = orc.training_loss(*out, rgb_gt.to(dev), depth_gt.to(dev))
            loss.backward()
            torch.cuda.synchronize()
        res[mode] = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
        if mode == "mma":
            sc = model._last_scratch
            zv = out[2].detach().cpu()
    listed = model.app_sample_count()[0]
    tiles = (listed + 127) // 128
    print(f"R={R} listed={listed} tiles={tiles} last tile on CTA {(tiles-1) % 148} as its tile #{(tiles-1)//148} with {listed - 128*(tiles-1)} live rows")
    for k in ("app_plane.0", "app_line.0", "basis_mat.weight", "renderModule.mlp.0.bias", "renderModule.mlp.2.bias", "renderModule.mlp.4.bias"):
        print(f"     {k:26s} mma-vs-ffma scaled err {scaled_err(res['mma'][k], res['ffma'][k]):.3e}")
    diff = (res["mma"]["app_plane.0"] - res["ffma"]["app_plane.0"]).abs().sum(1)[0]      # [H(y), W(x)]
    thr = 0.25 * float(diff.max())
    bad = torch.nonzero(diff > thr)
    print("     texels (y,x) with large app_plane.0 difference:", bad[:12].tolist())
    slots = sc["slots"][:listed].long().cpu()
    r, k = slots // S, slots % S
    z = zv[r, k]
    rc_ = rays.cpu()
    pts = rc_[r, :3] + rc_[r, 3:] * z[:, None]
    tx = (pts[:, 0] + 1.5) / 3.0 * 299.0
    ty = (pts[:, 1] + 1.5) / 3.0 * 299.0
    hit = torch.zeros(listed, dtype=torch.bool)
    for (by, bx) in bad[:50].tolist():
        hit |= ((tx - bx).abs() < 1.01) & ((ty - by).abs() < 1.01)
    e = torch.nonzero(hit).flatten()
    tiles_hit = sorted(set((e // 128).tolist()))
    print("     list entries on those texels:", e.numel(), "in tiles", tiles_hit[:20], "rows", sorted(set((e % 128).tolist()))[:40])
    print("     tile -> (CTA, tile# in CTA):", [(t, t % 148, t // 148) for t in tiles_hit[:10]])
