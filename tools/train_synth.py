#!/usr/bin/env python
"""Config-2 style training driver on synthetic data (BASELINE.json configs[1]; the loop of text2nerf_main.py:547-601).

No datasets or checkpoints exist offline, so the targets are TEACHER-RENDERED: a seeded fog field (the teacher) is rendered
from a ring of pin-hole cameras and a freshly initialised student is trained on those rays with the reference's iteration:

    ids = sampler.nextids(); rays, rgb, depth = batch
    rgb_map, depth_map, weights, z_vals = render(rays, is_train=True)            (text2nerf_main.py:556)
    loss = rgb MSE + 0.005 depth MSE + 1e3 TransMittanceLoss_mask                 (:563-575)
         + TV_loss_density(tvreg) * w_d + TV_loss_app(tvreg) * w_a                (:577-586)
    optimizer.zero_grad(); loss.backward(); optimizer.step()                      (:588-590)
    lr *= lr_factor                                                                (:600-601)

Two modes:
  parity   the SAME loop runs twice with identical seeds -- on the CPU oracle (plain tensor ops + torch.optim.Adam) and on
           the B200 path (fused data_loss + TV kernels + FusedAdam) -- on a small grid; the per-iteration PSNR trajectories
           must stay within 0.05 dB (north-star gate "PSNR within 0.05 dB of the reference on identical inputs").
  lego     lego-shaped run on the B200 path only: 300^3 target resolution reached coarse-to-fine (upsample_volume_grid),
           alpha-mask updates + shrink + ray filtering (the f3 maintenance kernels inside a real loop), 4096-ray batches;
           reports iterations/s, the PSNR curve and a held-out full-view render.

    python tools/train_synth.py parity [--iters 100] [--out profiles/r2_train_parity.json]
    python tools/train_synth.py lego   [--iters 3000] [--out profiles/r2_train_lego.json]
"""
import argparse
import contextlib
import io
import json
import math
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class TVLoss(torch.nn.Module):
    """utils.TVLoss (utils.py:488-504) as text2nerf_main.py:455 instantiates it."""

    def __init__(self, TVLoss_weight=1):
        super().__init__()
        self.TVLoss_weight = TVLoss_weight

    def forward(self, x):
        n_h, n_w = x[:, :, 1:, :].numel(), x[:, :, :, 1:].numel()
        h_tv = torch.pow(x[:, :, 1:, :] - x[:, :, :-1, :], 2).sum()
        w_tv = torch.pow(x[:, :, :, 1:] - x[:, :, :, :-1], 2).sum()
        return self.TVLoss_weight * 2 * (h_tv / n_h + w_tv / n_w) / x.shape[0]


class SimpleSampler:
    """renderer.SimpleSampler (renderer.py:14-26) with a private numpy generator."""

    def __init__(self, total, batch, seed):
        self.total, self.batch, self.curr, self.ids = total, batch, total, None
        self.rng = np.random.RandomState(seed)

    def nextids(self):
        self.curr += self.batch
        if self.curr + self.batch > self.total:
            self.ids = torch.LongTensor(self.rng.permutation(self.total))
            self.curr = 0
        return self.ids[self.curr:self.curr + self.batch]


def cam_dirs(H, W, focal):
    """Camera-space ray directions of a pin-hole view, pixel centres at +0.5, normalised (ray_utils.py:24-42, scene_gen.py:45)."""
    ys, xs = torch.meshgrid(torch.linspace(0, H - 1, H), torch.linspace(0, W - 1, W), indexing="ij")
    d = torch.stack([(xs + 0.5 - W / 2) / focal, (ys + 0.5 - H / 2) / focal, torch.ones_like(xs)], -1)
    return (d / d.norm(dim=-1, keepdim=True)).reshape(-1, 3)


def look_at_rays(eye, target, dirs):
    """[n, 6] rays of a camera at `eye` looking at `target` (rotation without renormalisation, ray_utils.py:66-87)."""
    eye, target = torch.as_tensor(eye, dtype=torch.float32), torch.as_tensor(target, dtype=torch.float32)
    fwd = (target - eye) / (target - eye).norm()
    right = torch.linalg.cross(torch.tensor([0.0, 1.0, 0.0]), fwd)
    right = right / right.norm()
    up = torch.linalg.cross(fwd, right)
    Rm = torch.stack([right, up, fwd], 1)                     # camera-to-world rotation
    return torch.cat([eye.expand(dirs.shape[0], 3), dirs @ Rm.T], -1).contiguous()


def psnr_from_mse(mse):
    return -10.0 * math.log(max(mse, 1e-20)) / math.log(10.0)           # text2nerf_main.py:596


# ------------------------------------------------------------------------------------------------------------------
# the loop on the CPU oracle
# ------------------------------------------------------------------------------------------------------------------
def train_oracle(spec, init, rays, rgbs, depths, iters, batch, S, seed, lr_xyz=0.02, lr_net=1e-3, tv_d=0.1, tv_a=0.01):
    from oracle import t2n_oracle as orc
    p = {k: v.clone().requires_grad_(True) for k, v in init.items()}
    groups = [{"params": [p[f"density_line.{i}"] for i in range(3)], "lr": lr_xyz},
              {"params": [p[f"density_plane.{i}"] for i in range(3)], "lr": lr_xyz},
              {"params": [p[f"app_line.{i}"] for i in range(3)], "lr": lr_xyz},
              {"params": [p[f"app_plane.{i}"] for i in range(3)], "lr": lr_xyz},
              {"params": [p["basis_mat.weight"]], "lr": lr_net},
              {"params": [p[k] for k in p if k.startswith("renderModule")], "lr": lr_net}]   # tensoRF.py:164-170
    opt = torch.optim.Adam(groups, betas=(0.9, 0.99))
    lr_factor = 0.1 ** (1.0 / iters)
    sampler = SimpleSampler(rays.shape[0], batch, seed)
    torch.manual_seed(seed)
    curve = []
    for it in range(iters):
        ids = sampler.nextids()
        jitter = torch.rand(batch, 1)                                   # tensorBase.py:313-317
        out = orc.render(spec, p, rays[ids], S, True, True, jitter)
        loss = orc.training_loss(*out, rgbs[ids], depths[ids])
        loss = loss + sum(orc.tv_plane(p[f"density_plane.{i}"]) for i in range(3)) * 1e-2 * tv_d
        loss = loss + sum(orc.tv_plane(p[f"app_plane.{i}"]) for i in range(3)) * 1e-2 * tv_a
        opt.zero_grad()
        loss.backward()
        opt.step()
        for g in opt.param_groups:
            g["lr"] = g["lr"] * lr_factor
        curve.append(psnr_from_mse(float(((out[0].detach() - rgbs[ids]) ** 2).mean())))
    return curve, {k: v.detach() for k, v in p.items()}


# ------------------------------------------------------------------------------------------------------------------
# the loop on the B200 path
# ------------------------------------------------------------------------------------------------------------------
def build_student(spec, init, dev):
    from text2nerf_b200 import TensorVMSplit
    with contextlib.redirect_stdout(io.StringIO()):
        m = TensorVMSplit(spec.aabb_t().to(dev), list(spec.grid), dev, density_n_comp=list(spec.density_n_comp),
                          appearance_n_comp=list(spec.app_n_comp), app_dim=spec.app_dim, near_far=list(spec.near_far),
                          shadingMode=spec.shading, alphaMask_thres=0.001, density_shift=spec.density_shift,
                          distance_scale=spec.distance_scale, pos_pe=spec.pos_pe, view_pe=spec.view_pe, fea_pe=spec.fea_pe,
                          featureC=spec.featureC, step_ratio=spec.step_ratio, fea2denseAct=spec.act)
    m.load_state_dict({k: v.to(dev) for k, v in init.items()})
    return m


def train_b200(spec, init, rays, rgbs, depths, iters, batch, S, seed, dev, lr_xyz=0.02, lr_net=1e-3, tv_d=0.1, tv_a=0.01):
    from text2nerf_b200.optim import FusedAdam
    model = build_student(spec, init, dev)
    opt = FusedAdam(model.get_optparam_groups(lr_xyz, lr_net), betas=(0.9, 0.99))
    tvreg = TVLoss()
    lr_factor = 0.1 ** (1.0 / iters)
    sampler = SimpleSampler(rays.shape[0], batch, seed)
    rays_d, rgbs_d, depths_d = rays.to(dev), rgbs.to(dev), depths.to(dev)
    torch.manual_seed(seed)
    curve = []
    for it in range(iters):
        ids = sampler.nextids().to(dev)
        loss, l_rgb, _, _ = model.data_loss(rays_d[ids], rgbs_d[ids], depths_d[ids], white_bg=True, N_samples=S,
                                            return_terms=True)          # draws torch.rand(batch, 1) like the reference
        total = loss + model.TV_loss_density(tvreg) * tv_d + model.TV_loss_app(tvreg) * tv_a
        opt.zero_grad()
        total.backward()
        opt.step()
        for g in opt.param_groups:
            g["lr"] = g["lr"] * lr_factor
        curve.append(l_rgb)                                             # device scalars: read back once at the end
    curve = [psnr_from_mse(float(v)) for v in torch.stack(curve).cpu()]
    return curve, model


def teacher_targets(spec, teacher, rays, S, chunk=4096):
    """rgb / depth of the teacher field along the training rays (evaluation render of the CPU oracle)."""
    from oracle import t2n_oracle as orc
    rgb, depth = [], []
    with torch.no_grad():
        for s in range(0, rays.shape[0], chunk):
            o = orc.render(spec, teacher, rays[s:s + chunk], S, False, True, None)
            rgb.append(o[0]); depth.append(o[1])
    return torch.cat(rgb), torch.cat(depth)


def run_parity(iters=100, batch=256, grid=32, seed=3, dev=None):
    from oracle import t2n_oracle as orc
    dev = dev or torch.device("cuda:0")
    spec = orc.FieldSpec(aabb=[[-1.5, -1.5, 2.5], [1.5, 1.5, 5.5]], grid=[grid] * 3, near_far=[2.0, 6.0], step_ratio=0.5)
    S = orc.derive_step(spec)[1]
    teacher = orc.init_params(spec, seed=seed, density_gain=12.0, app_gain=3.0)
    student = orc.init_params(spec, seed=seed + 1, density_gain=6.0, app_gain=1.0)
    # cameras in front of the box, which sits at z in [2.5, 5.5]: the reference drops world z <= 2 at evaluation
    # (tensorBase.py:459-462), so the teacher views are taken from around the origin
    dirs = cam_dirs(32, 32, 44.0)
    eyes = 0.4 * torch.randn(4, 3, generator=torch.Generator().manual_seed(seed))
    rays = torch.cat([look_at_rays(e, [0.0, 0.0, 4.0], dirs) for e in eyes])
    rgbs, depths = teacher_targets(spec, teacher, rays, S)
    t0 = time.time()
    curve_ref, p_ref = train_oracle(spec, student, rays, rgbs, depths, iters, batch, S, seed)
    t_ref = time.time() - t0
    t0 = time.time()
    curve_gpu, model = train_b200(spec, student, rays, rgbs, depths, iters, batch, S, seed, dev)
    torch.cuda.synchronize()
    t_gpu = time.time() - t0
    diff = [abs(a - b) for a, b in zip(curve_ref, curve_gpu)]
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    drift = max(float((sd[k] - p_ref[k]).abs().max() / p_ref[k].abs().max().clamp_min(1e-30)) for k in p_ref)
    return {"mode": "parity", "grid": grid, "iters": iters, "batch": batch, "S": S, "rays": int(rays.shape[0]),
            "psnr_first": [curve_ref[0], curve_gpu[0]], "psnr_last": [curve_ref[-1], curve_gpu[-1]],
            "psnr_max_abs_diff_db": max(diff), "psnr_mean_abs_diff_db": sum(diff) / len(diff),
            "param_max_scaled_drift": drift, "gate_db": 0.05, "ok": bool(max(diff) <= 0.05),
            "oracle_seconds": t_ref, "b200_seconds": t_gpu, "curve_oracle": curve_ref, "curve_b200": curve_gpu}


def run_lego(iters=3000, batch=4096, dev=None, seed=5, n_views=24, hw=200):
    """Coarse-to-fine lego-shaped run: 128^3 -> 300^3 by upsample_volume_grid at 5 points of the schedule, alpha-mask update
    + shrink after the first stage, ray filtering with the mask, TV regularisers, FusedAdam with the learning-rate reset
    TensoRF applies after every upsampling (lr_upsample_reset)."""
    from oracle import t2n_oracle as orc
    from text2nerf_b200 import OctreeRender_trilinear_fast
    from text2nerf_b200.optim import FusedAdam
    dev = dev or torch.device("cuda:0")
    aabb = [[-1.5, -1.5, 2.5], [1.5, 1.5, 5.5]]
    final = 300
    spec_t = orc.FieldSpec(aabb=aabb, grid=[final] * 3, near_far=[2.0, 6.0], step_ratio=0.5)
    teacher_p = orc.init_params(spec_t, seed=seed, density_gain=10.8, app_gain=3.0)
    teacher = build_student(spec_t, teacher_p, dev)
    S_t = teacher.nSamples
    # cameras in front of the box (world z of every sample > 2: the eval-only filter of tensorBase.py:459-462)
    rays = []
    dirs = cam_dirs(hw, hw, 1.39 * hw)
    for v in range(n_views + 1):
        eye = [1.2 * math.cos(2 * math.pi * (v + 0.5 * (v == n_views)) / n_views),
               0.9 * math.sin(2 * math.pi * (v + 0.5 * (v == n_views)) / n_views), 0.0]
        rays.append(look_at_rays(eye, [0.0, 0.0, 4.0], dirs))
    train_rays, test_rays = torch.cat(rays[:-1]).contiguous(), rays[-1].contiguous()
    with torch.no_grad():
        rgb_t, _, dep_t, _, _ = OctreeRender_trilinear_fast(train_rays, teacher, chunk=1 << 18, N_samples=S_t, white_bg=True,
                                                            is_train=False, device=dev)
        rgb_test, _, _, _, _ = OctreeRender_trilinear_fast(test_rays, teacher, chunk=1 << 18, N_samples=S_t, white_bg=True,
                                                           is_train=False, device=dev)
    del teacher
    reso0 = 128
    spec_s = orc.FieldSpec(aabb=aabb, grid=[reso0] * 3, near_far=[2.0, 6.0], step_ratio=0.5)
    model = build_student(spec_s, orc.init_params(spec_s, seed=seed + 1, density_gain=1.0, app_gain=1.0), dev)
    with contextlib.redirect_stdout(io.StringIO()):
        rays_f, rgb_f, dep_f = model.filtering_rays(train_rays, rgb_t.cpu(), all_depth=dep_t.cpu(), bbox_only=True)
    rays_d, rgbs_d, deps_d = rays_f.to(dev), rgb_f.to(dev), dep_f.to(dev)
    ups = [int(iters * f) for f in (0.15, 0.25, 0.35, 0.45, 0.55)]
    n_vox = [int(round(math.exp(x))) for x in np.linspace(math.log(reso0 ** 3), math.log(final ** 3), len(ups) + 1)][1:]
    mask_at = [ups[0], ups[2]]
    lr_xyz, lr_net = 0.02, 1e-3
    opt = FusedAdam(model.get_optparam_groups(lr_xyz, lr_net), betas=(0.9, 0.99))
    lr_factor = 0.1 ** (1.0 / iters)
    tvreg = TVLoss()
    sampler = SimpleSampler(rays_d.shape[0], batch, seed)
    torch.manual_seed(seed)
    curve, events = [], []
    torch.cuda.synchronize()
    t0 = time.time()
    t_maint = 0.0
    for it in range(iters):
        ids = sampler.nextids().to(dev)
        S = min(model.nSamples, S_t)
        loss, l_rgb, _, _ = model.data_loss(rays_d[ids], rgbs_d[ids], deps_d[ids], white_bg=True, N_samples=S, return_terms=True)
        total = loss + model.TV_loss_density(tvreg) * 0.1 + model.TV_loss_app(tvreg) * 0.01
        opt.zero_grad()
        total.backward()
        opt.step()
        for gp in opt.param_groups:
            gp["lr"] = gp["lr"] * lr_factor
        curve.append(l_rgb)
        if it in mask_at or it in ups:
            torch.cuda.synchronize()
            tm = time.time()
            with contextlib.redirect_stdout(io.StringIO()):
                if it in mask_at:
                    g3 = [int(v) for v in model.gridSize.tolist()]
                    new_aabb = model.updateAlphaMask(tuple(g3))
                    if it == mask_at[0]:
                        model.shrink(new_aabb)
                    if it == mask_at[-1]:
                        r2, c2, d2 = model.filtering_rays(rays_d.cpu(), rgbs_d.cpu(), all_depth=deps_d.cpu(), N_samples=S)
                        rays_d, rgbs_d, deps_d = r2.to(dev), c2.to(dev), d2.to(dev)
                        sampler = SimpleSampler(rays_d.shape[0], batch, seed + it)
                    events.append({"iter": it, "event": "alpha mask", "grid": g3, "rays_left": int(rays_d.shape[0]),
                                   "occupied_pct": float(model.alphaMask.alpha_volume.mean() * 100)})
                if it in ups:
                    nv = n_vox[ups.index(it)]
                    size = (model.aabb[1] - model.aabb[0]).cpu()
                    vs = float((size.prod() / nv) ** (1 / 3))
                    res = [int(v) for v in (size / vs).long().tolist()]          # utils.N_to_reso (utils.py:292-296)
                    model.upsample_volume_grid(res)
                    scale = 0.1 ** (it / iters)                                   # lr_upsample_reset = 0: continue the decay
                    opt = FusedAdam(model.get_optparam_groups(lr_xyz * scale, lr_net * scale), betas=(0.9, 0.99))
                    events.append({"iter": it, "event": "upsample", "grid": res, "nSamples": model.nSamples})
            torch.cuda.synchronize()
            t_maint += time.time() - tm
    torch.cuda.synchronize()
    t_total = time.time() - t0
    curve = [psnr_from_mse(float(v)) for v in torch.stack(curve).cpu()]
    with torch.no_grad():
        rgb_o, _, _, _, _ = OctreeRender_trilinear_fast(test_rays, model, chunk=1 << 18, N_samples=S_t, white_bg=True,
                                                        is_train=False, device=dev)
    test_psnr = psnr_from_mse(float(((rgb_o - rgb_test) ** 2).mean()))
    k = max(1, iters // 20)
    return {"mode": "lego", "iters": iters, "batch": batch, "train_rays": int(train_rays.shape[0]), "views": n_views,
            "view_hw": hw, "final_grid": [int(v) for v in model.gridSize.tolist()], "final_nSamples": model.nSamples,
            "seconds_total": t_total, "seconds_maintenance": t_maint, "iterations_per_s": iters / (t_total - t_maint),
            "train_psnr_first": sum(curve[:k]) / k, "train_psnr_last": sum(curve[-k:]) / k, "heldout_view_psnr": test_psnr,
            "events": events, "curve_every_%d" % k: curve[::k]}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["parity", "lego"])
    ap.add_argument("--iters", type=int, default=None)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    res = run_parity(a.iters or 100) if a.mode == "parity" else run_lego(a.iters or 3000)
    txt = json.dumps(res)
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        open(a.out, "w").write(txt + "\n")
    brief = {k: v for k, v in res.items() if not k.startswith("curve")}
    print(json.dumps(brief))
