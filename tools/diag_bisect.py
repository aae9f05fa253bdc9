"""Diagnostic: bisect the ray whose samples make the tensor-core backward differ from the FFMA backward."""
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
R = 512
g = torch.Generator().manual_seed(9)
px = torch.rand(R, 2, generator=g) * 800.0
d = torch.cat([(px - 400.0) / 1111.1, torch.ones(R, 1)], -1)
rays_all = torch.cat([torch.zeros(R, 3), d / d.norm(dim=-1, keepdim=True)], -1).contiguous().to(dev)
jitter_all = torch.rand(R, 1, generator=g)
rgb_gt_all = torch.rand(R, 3, generator=g)
depth_gt_all = 2.0 + 4.0 * torch.rand(R, generator=g)


def grads(idx, mode):
    if mode == "ffma":
        os.environ["T2N_BWD_FFMA"] = "1"
    else:
        os.environ.pop("T2N_BWD_FFMA", None)
    rays, jit = rays_all[idx], jitter_all[idx.cpu()]
    out = None
    for rep in range(2):
        model.zero_grad()
        out = render_with_jitter(model, rays, jit, True, True, S)
        # same per-ray upstream gradients whatever the subset: sum (not mean) losses
        loss = ((out[0] - rgb_gt_all[idx.cpu()].to(dev)) ** 2).sum() / 1536.0 + 0.005 * ((out[1] - depth_gt_all[idx.cpu()].to(dev)) ** 2).sum() / 512.0
        loss.backward()
        torch.cuda.synchronize()
    return {k: p.grad.detach().clone() for k, p in model.named_parameters()}, out


def err(idx):
    a, _ = grads(idx, "mma")
    b, _ = grads(idx, "ffma")
    return float((a["app_plane.0"] - b["app_plane.0"]).abs().max()), a, b


idx = torch.arange(R, device=dev)
e0, _, _ = err(idx)
print("full batch abs err", e0)
while idx.numel() > 1:
    h = idx.numel() // 2
    left, right = idx[:h], idx[h:]
    el, _, _ = err(left)
    er, _, _ = err(right)
    print(f"n={idx.numel()} left {el:.3e} right {er:.3e}")
    idx = left if el >= er else right
print("ray", int(idx[0]), rays_all[idx].cpu().tolist(), "jitter", float(jitter_all[int(idx[0])]))
e1, a, b = err(idx)
print("single ray abs err", e1, "max |g| app_plane.0", float(b["app_plane.0"].abs().max()))
_, out = grads(idx, "mma")
w = out[3][0].detach().cpu()
z = out[2][0].cpu()
listed = torch.nonzero(w > 1e-4).flatten()
print("listed samples", listed.numel(), "k range", int(listed.min()), int(listed.max()))
# per z-texel difference of app_line.0 (line 0 runs along z)
dl = (a["app_line.0"] - b["app_line.0"]).abs().sum(1)[0, :, 0].cpu()
gl = b["app_line.0"].abs().sum(1)[0, :, 0].cpu()
nz = torch.nonzero(gl > 0).flatten()
print("z texels touched", int(nz.min()), int(nz.max()))
for t in nz.tolist():
    print(f"   z texel {t}: |ffma| {float(gl[t]):.3e} |diff| {float(dl[t]):.3e} rel {float(dl[t] / gl[t]):.2e}")
rr = rays_all[idx][0].cpu()
for k in listed.tolist():
    p = rr[:3] + rr[3:] * z[k]
    print(f"   k={k} z={float(z[k]):.5f} w={float(w[k]):.5f} texel z={(float(p[2]) - 2.5) / 3 * 299:.3f} x={(float(p[0]) + 1.5) / 3 * 299:.3f} y={(float(p[1]) + 1.5) / 3 * 299:.3f}")
for k in ("renderModule.mlp.0.bias", "renderModule.mlp.2.bias", "renderModule.mlp.4.bias", "basis_mat.weight"):
    print(k, scaled_err(a[k], b[k]))
