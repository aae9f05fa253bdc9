"""Generate the golden fixtures in this directory by running the UNMODIFIED reference
(eckertzhang/Text2NeRF) on CPU.  Only runs where the read-only reference checkout exists
(the authoring container, /root/reference); the .npz files it writes are committed and are
what travels to the GPU box.

    python tests/golden/make_golden.py [--ref /root/reference]

Every case stores: the FieldSpec (json), the parameters (reference state-dict keys), the
rays, the per-ray training jitter the reference drew from the CPU RNG (replayed from the
same seed), the four forward outputs, and the gradients of the Text2NeRF data loss
(text2nerf_main.py:563-575) w.r.t. every parameter.
"""
import argparse
import io
import json
import os
import sys
import types
from contextlib import redirect_stdout
from dataclasses import asdict

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import t2n_oracle as orc  # noqa: E402


def stub_optional_modules():
    """renderer.py / ray_utils.py import packages that are absent here; none is used by
    the functions we call except kornia.create_meshgrid (ray_utils.py:34)."""
    def create_meshgrid(H, W, normalized_coordinates=False):
        xs = torch.linspace(0, W - 1, W)
        ys = torch.linspace(0, H - 1, H)
        gy, gx = torch.meshgrid(ys, xs, indexing="ij")
        return torch.stack([gx, gy], -1)[None]
    for name in ("kornia", "imageio", "imageio.v2", "statsmodels", "statsmodels.api", "skimage",
                 "skimage.io", "skimage.measure", "plyfile", "lpips"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["kornia"].create_meshgrid = create_meshgrid
    sys.modules["plyfile"].PlyData = object
    sys.modules["plyfile"].PlyElement = object


def build_reference(ref_mod, spec: orc.FieldSpec, params):
    with redirect_stdout(io.StringIO()):
        m = ref_mod.TensorVMSplit(
            spec.aabb_t(), list(spec.grid), "cpu",
            density_n_comp=list(spec.density_n_comp), appearance_n_comp=list(spec.app_n_comp),
            app_dim=spec.app_dim, near_far=list(spec.near_far), shadingMode=spec.shading,
            alphaMask_thres=0.001, density_shift=spec.density_shift, distance_scale=spec.distance_scale,
            pos_pe=spec.pos_pe, view_pe=spec.view_pe, fea_pe=spec.fea_pe, featureC=spec.featureC,
            step_ratio=spec.step_ratio, fea2denseAct=spec.act)
    m.load_state_dict(params)
    return m


def pinhole_rays(n, origin, spread, seed, unit=True):
    g = torch.Generator().manual_seed(seed)
    d = torch.cat([spread * (torch.rand(n, 2, generator=g) * 2 - 1), torch.ones(n, 1)], -1)
    if unit:
        d = d / d.norm(dim=-1, keepdim=True)
    o = torch.tensor(origin, dtype=torch.float32).expand(n, 3) + 0.05 * torch.randn(n, 3, generator=g)
    return torch.cat([o, d], -1).contiguous()


def spec_json(spec):
    d = asdict(spec)
    d["dtype"] = "float32"
    return json.dumps(d)


def run_case(ref_mod, name, spec, params, rays, is_train, white_bg, n_samples, seed, alpha=None,
             with_grads=True):
    model = build_reference(ref_mod, spec, params)
    if alpha is not None:
        vol, maabb = alpha
        model.alphaMask = ref_mod.AlphaGridMask("cpu", maabb, vol[0, 0])
    R = rays.shape[0]
    torch.manual_seed(seed)
    jitter = torch.rand(R, 1) if is_train else None      # replay of tensorBase.py:316
    # tensorBase.py:497: `white_bg or (is_train and torch.rand((1,))<0.5)` draws one more CPU
    # uniform only when white_bg is False; the effective flag is what the oracle takes
    eff_white = white_bg or (is_train and bool(torch.rand((1,)) < 0.5))
    torch.manual_seed(seed)
    rgb, depth, z, w = model(rays, is_train=is_train, white_bg=white_bg, ndc_ray=0, N_samples=n_samples)

    # the oracle must reproduce the reference bit for bit on the forward outputs
    o_rgb, o_depth, o_z, o_w = orc.render(spec, params, rays, n_samples, is_train, eff_white, jitter, alpha)
    for a, b, what in ((rgb, o_rgb, "rgb"), (depth, o_depth, "depth"), (z, o_z, "z"), (w, o_w, "weight")):
        assert torch.equal(a.detach(), b.detach()), f"{name}: oracle != reference on {what}"

    out = {"spec": np.array(spec_json(spec)), "rays": rays.numpy(), "is_train": np.array(is_train),
           "white_bg": np.array(white_bg), "white_bg_effective": np.array(eff_white), "n_samples": np.array(n_samples),
           "rgb_map": rgb.detach().numpy(), "depth_map": depth.detach().numpy(),
           "z_vals": z.detach().numpy(), "weight": w.detach().numpy()}
    if jitter is not None:
        out["jitter"] = jitter.numpy()
    if alpha is not None:
        out["alpha_volume"] = alpha[0].numpy()
        out["alpha_aabb"] = alpha[1].numpy()
    for k, v in params.items():
        out["param/" + k] = v.numpy()
    if with_grads:
        g = torch.Generator().manual_seed(seed + 1)
        rgb_gt = torch.rand(R, 3, generator=g)
        depth_gt = spec.near_far[0] + (spec.near_far[1] - spec.near_far[0]) * torch.rand(R, generator=g)
        loss = orc.training_loss(rgb, depth, z, w, rgb_gt, depth_gt)
        model.zero_grad()
        loss.backward()
        out["rgb_gt"], out["depth_gt"], out["loss"] = rgb_gt.numpy(), depth_gt.numpy(), loss.detach().numpy()
        for k, p in model.named_parameters():
            out["grad/" + k] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    n_app = int((w > spec.weight_thres).sum())
    print(f"{name}: R={R} S={z.shape[1]} app={n_app} acc_mean={float(w.detach().sum(-1).mean()):.3f} "
          f"rgb_mean={float(rgb.mean()):.3f}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    args = ap.parse_args()
    sys.path.insert(0, args.ref)
    stub_optional_modules()
    import models.tensoRF as ref_mod            # noqa: E402  (the unmodified reference)
    from dataLoader import ray_utils as ref_rays  # noqa: E402

    torch.set_num_threads(1)

    # ---- A: the configured Text2NeRF field (configs/text2nerf_scenes.txt), small non-cubic grid
    specA = orc.FieldSpec(aabb=[[-8, -8, -8], [8, 8, 8]], grid=[18, 22, 26], near_far=[0.5, 8.0],
                          step_ratio=1.0, shading="MLP_Fea_noview")
    pA = orc.init_params(specA, seed=11, density_gain=10.8, app_gain=3.0)
    raysA = pinhole_rays(80, [0.3, -0.2, 0.1], 0.6, seed=5)
    nA = orc.derive_step(specA)[1]
    run_case(ref_mod, "t2n_noview_train", specA, pA, raysA, True, True, nA // 2, seed=101)
    run_case(ref_mod, "t2n_noview_eval", specA, pA, raysA, False, True, -1, seed=102, with_grads=False)
    run_case(ref_mod, "t2n_noview_train_blackbg", specA, pA, raysA, True, False, nA // 2, seed=103)

    # ---- B: lego-shaped box pushed to z in [2.5,5.5], other shading heads, small nets
    def specB(mode, **kw):
        base = dict(aabb=[[-1.5, -1.2, 2.5], [1.5, 1.8, 5.5]], grid=[24, 20, 28], near_far=[2.0, 6.0],
                    step_ratio=0.5, shading=mode, featureC=32, density_n_comp=(4, 8, 12),
                    app_n_comp=(8, 12, 4), pos_pe=6, view_pe=6, fea_pe=2)
        base.update(kw)
        return orc.FieldSpec(**base)

    raysB = pinhole_rays(64, [0.0, 0.1, 0.0], 0.35, seed=7)
    # edge cases: axis-parallel ray (d components exactly 0), a ray that misses the box, origin inside box
    raysB[0] = torch.tensor([0.2, 0.3, 0.0, 0.0, 0.0, 1.0])
    raysB[1] = torch.tensor([5.0, 5.0, 0.0, 0.0, 0.0, 1.0])
    raysB[2] = torch.tensor([0.1, 0.2, 3.0, 0.3, -0.2, 0.93])
    raysB[3] = torch.tensor([0.0, 0.0, 0.0, 1.0, 0.0, 0.0])
    # MLP_PE is not a working head in the reference: its first Linear is sized with 3 extra `pts`
    # columns (tensorBase.py:115) that forward never concatenates (:124-130) -> shape error; no golden.
    for mode, kw in (("MLP_Fea", dict(view_pe=2)), ("MLP", {}), ("SH", {}), ("RGB", dict(app_dim=3))):
        s = specB(mode, **kw)
        p = orc.init_params(s, seed=21, density_gain=22.0, app_gain=3.0)
        nB = orc.derive_step(s)[1]
        run_case(ref_mod, f"lego_{mode.lower()}_train", s, p, raysB, True, True, nB // 3, seed=201)
        run_case(ref_mod, f"lego_{mode.lower()}_eval", s, p, raysB, False, True, nB // 3, seed=202,
                 with_grads=False)

    # ---- C: relu activation + alpha mask volume (TensoRF-style driver surface)
    sC = specB("MLP_Fea_noview", act="relu", fea_pe=2)
    pC = orc.init_params(sC, seed=31, density_gain=6.0, app_gain=3.0)
    g = torch.Generator().manual_seed(3)
    vol = (torch.rand(1, 1, 9, 11, 13, generator=g) > 0.45).float()
    maabb = torch.tensor([[-1.4, -1.1, 2.6], [1.4, 1.7, 5.4]])
    nC = orc.derive_step(sC)[1]
    run_case(ref_mod, "lego_relu_alphamask_train", sC, pC, raysB, True, True, nC // 3, seed=301,
             alpha=(vol, maabb))

    # ---- D: ray generation (dataLoader/ray_utils.py:24-42, 66-87)
    H, W, focal = 6, 8, [7.5, 7.25]
    dirs = ref_rays.get_ray_directions(H, W, focal)
    dirs_n = dirs / torch.norm(dirs, dim=-1, keepdim=True)       # scene_gen.py:45
    g = torch.Generator().manual_seed(9)
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
    c2w = torch.cat([q, torch.randn(3, 1, generator=g)], 1)
    ro, rd = ref_rays.get_rays(dirs_n, c2w)
    o_dirs = orc.pixel_directions(H, W, focal)
    assert torch.equal(o_dirs, dirs)
    o_ro, o_rd = orc.camera_rays(o_dirs / torch.norm(o_dirs, dim=-1, keepdim=True), c2w)
    assert torch.equal(o_ro, ro) and torch.equal(o_rd, rd)
    np.savez_compressed(os.path.join(HERE, "get_rays.npz"), H=H, W=W, focal=np.array(focal), c2w=c2w.numpy(),
                        directions=dirs.numpy(), rays_o=ro.numpy(), rays_d=rd.numpy())
    print("get_rays: ok")


if __name__ == "__main__":
    main()
