"""Diagnostic: per-parameter gradient error at the bench shape (300^3, S=1036, 512 rays), composed vs fused loss."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import t2n_oracle as orc
from helpers import build_model, render_with_jitter, fused_loss_with_jitter, scaled_err, cosine

dev = torch.device("cuda:0")
spec = orc.FieldSpec(aabb=[[-1.5, -1.5, 2.5], [1.5, 1.5, 5.5]], grid=[300, 300, 300], near_far=[2.0, 6.0], step_ratio=0.5)
params = orc.init_params(spec, seed=0, density_gain=10.8, app_gain=1.0)
S = orc.derive_step(spec)[1]
g = torch.Generator().manual_seed(9)
R = int(os.environ.get("R", 512))
px = torch.rand(R, 2, generator=g) * 800.0
d = torch.cat([(px - 400.0) / 1111.1, torch.ones(R, 1)], -1)
rays = torch.cat([torch.zeros(R, 3), d / d.norm(dim=-1, keepdim=True)], -1).contiguous()
jitter = torch.rand(R, 1, generator=g)
rgb_gt = torch.rand(R, 3, generator=g)
depth_gt = 2.0 + 4.0 * torch.rand(R, generator=g)
p_ref = {k: v.clone().requires_grad_(True) for k, v in params.items()}
ref_t = orc.render(spec, p_ref, rays, S, True, True, jitter, None, keep=True)
aux = ref_t[4]
loss_ref = orc.training_loss(*ref_t[:4], rgb_gt, depth_gt)
loss_ref.backward()
model = build_model(spec, params, dev)
for mode in ("composed", "fused", "composed_ffma"):
    if mode == "composed_ffma":
        os.environ["T2N_DECODER"] = "ffma"
    for rep in range(2):
        model.zero_grad()
        if mode.startswith("composed"):
            out = render_with_jitter(model, rays.to(dev), jitter, True, True, S)
            loss = orc.training_loss(*out, rgb_gt.to(dev), depth_gt.to(dev))
        else:
            loss = fused_loss_with_jitter(model, rays.to(dev), jitter, True, S, rgb_gt, depth_gt)[0]
        loss.backward()
        torch.cuda.synchronize()
    print(mode, "loss rel", abs(float(loss) - float(loss_ref)) / abs(float(loss_ref)), "listed", model.app_sample_count())
    for k, p in model.named_parameters():
        gr = p_ref[k].grad
        e = scaled_err(p.grad, gr)
        flag = " <<<<" if e > 2e-4 else ""
        print(f"   {k:28s} scaled_err {e:.3e} cos-1 {cosine(p.grad, gr)-1:.2e} max|g| {float(gr.abs().max()):.3e}{flag}")
        if e > 2e-4 and p.dim() == 4:
            diff = (p.grad.cpu() - gr).abs()
            idx = torch.nonzero(diff > 0.5 * diff.max())
            for i in idx[:6]:
                i = tuple(i.tolist())
                print("      at", i, "ours", float(p.grad.cpu()[i]), "ref", float(gr[i]))
if mode.startswith("composed"):
    w = out[3].detach().cpu()
    print("weight max abs err", float((w - ref_t[3].detach()).abs().max()), "flips", int(((w > 1e-4) != aux["app_mask"]).sum()))
    print("rgb err", float((out[0].detach().cpu() - ref_t[0].detach()).abs().max()))
    cl = (ref_t[0].detach() >= 1.0) | (ref_t[0].detach() <= 0.0)
    print("reference rays at the clamp:", int(cl.any(-1).sum()))
