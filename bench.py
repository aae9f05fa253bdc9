#!/usr/bin/env python
"""bench.py -- throughput of the TensoRF ray-marching hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    (N > 1: launched by torch.distributed.run, one rank per GPU)

Workload (BASELINE.json configs[1], "lego 800x800"): TensorVMSplit 300^3 (16/48 components,
MLP_Fea_noview 351->128->128->3), box +-1.5 pushed to z in [2.5, 5.5] (the reference's eval path
drops world z <= 2, tensorBase.py:459-462), step_ratio 0.5 -> 1036 samples per ray, one 800x800
pin-hole view = 640,000 rays per step, synthetic "fog" field (seeded init, density factors x10.8,
SURVEY.md 8d).  Data is synthetic and weights are random: no datasets exist offline.

One step = one full-view forward render (eval, no_grad) through the fused kernels.
  value      Mrays/s of that step with rays already resident in HBM (whole job, all ranks)
  e2e        the same view through the reference-facing API OctreeRender_trilinear_fast with the
             rays in pinned HOST memory (H2D inside the timed region) and rgb/depth maps read
             back to the host (D2H), as renderer.evaluation does per view (renderer.py:89-93)
  fwd_bwd    second timed region: forward+backward of the Text2NeRF data loss
             (text2nerf_main.py:563-575) on 4096-ray training batches; with N > 1 ranks every step
             ends with ONE NCCL all-reduce of the flat gradient buffer
  roofline   dominant kernel of the forward: algorithmic bytes / CUDA-event time
  cpu_baseline  the oracle (CPU port of the reference, same ATen ops) on a bounded ray sample
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H = W = 800
FOCAL = 1111.1
GRID = [300, 300, 300]
AABB = [[-1.5, -1.5, 2.5], [1.5, 1.5, 5.5]]
NEAR_FAR = [2.0, 6.0]
STEP_RATIO = 0.5
TRAIN_BATCH = 4096
TRAIN_BATCHES_PER_STEP = 8
CPU_RAYS_FWD = 2048
CPU_RAYS_BWD = 512
METRIC = "Mrays/s forward (eval) full-view render at 800x800, lego-shaped TensorVMSplit 300^3"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    return ap.parse_args()


# ---- the synthetic workload is generated here, by plain tensor code: the product arm never imports oracle/ ----------
N_SIGMA, N_APP, APP_DIM, FEATURE_C, MLP_IN = [16, 16, 16], [48, 48, 48], 27, 128, 27 + 2 * 6 * 27   # MLP_Fea_noview, fea_pe 6
MAT_AXES, VEC_AXIS = ((0, 1), (0, 2), (1, 2)), (2, 1, 0)          # models/tensorBase.py:190-191


def make_params(seed=0, density_gain=10.8, app_gain=1.0):
    """Seeded "fog" field with the reference's state-dict keys and shapes (models/tensoRF.py:144-160, tensorBase.py:94-99):
    0.1 * randn factors, the density factors scaled so that ~10 % of the box is occupied."""
    import math
    g = torch.Generator().manual_seed(seed)
    p = {}

    def randn(*shape):
        return torch.randn(*shape, generator=g, dtype=torch.float32)

    for i in range(3):
        (a0, a1), v = MAT_AXES[i], VEC_AXIS[i]
        p[f"density_plane.{i}"] = 0.1 * density_gain * randn(1, N_SIGMA[i], GRID[a1], GRID[a0])
        p[f"density_line.{i}"] = 0.1 * density_gain * randn(1, N_SIGMA[i], GRID[v], 1)
        p[f"app_plane.{i}"] = 0.1 * app_gain * randn(1, N_APP[i], GRID[a1], GRID[a0])
        p[f"app_line.{i}"] = 0.1 * app_gain * randn(1, N_APP[i], GRID[v], 1)
    p["basis_mat.weight"] = randn(APP_DIM, sum(N_APP)) / math.sqrt(sum(N_APP))
    p["renderModule.mlp.0.weight"] = randn(FEATURE_C, MLP_IN) / math.sqrt(MLP_IN)
    p["renderModule.mlp.0.bias"] = 0.1 * randn(FEATURE_C)
    p["renderModule.mlp.2.weight"] = randn(FEATURE_C, FEATURE_C) / math.sqrt(FEATURE_C)
    p["renderModule.mlp.2.bias"] = 0.1 * randn(FEATURE_C)
    p["renderModule.mlp.4.weight"] = randn(3, FEATURE_C) / math.sqrt(FEATURE_C)
    p["renderModule.mlp.4.bias"] = torch.zeros(3)
    return p


def composed_loss(rgb_map, depth_map, z_vals, weight, rgb_gt, depth_gt, w_depth=0.005, w_trans=1e3, delta=0.1):
    """The data terms as the reference's loop composes them from tensor ops (text2nerf_main.py:563-575,
    utils.TransMittanceLoss_mask utils.py:67-80): the comparison point of the fused TensorBase.data_loss."""
    l_rgb = torch.mean((rgb_map - rgb_gt) ** 2)
    l_depth = torch.mean((depth_map - depth_gt) ** 2)
    mean_w = torch.mean(weight * ((z_vals - depth_gt[:, None] + delta) < 0), dim=1)
    return l_rgb + w_depth * l_depth + w_trans * torch.mean(mean_w ** 2)


class TVLoss(torch.nn.Module):
    """utils.TVLoss (utils.py:488-504) as text2nerf_main.py:455 instantiates it (`tvreg`)."""

    def __init__(self, TVLoss_weight=1):
        super().__init__()
        self.TVLoss_weight = TVLoss_weight

    def forward(self, x):
        n_h, n_w = x[:, :, 1:, :].numel(), x[:, :, :, 1:].numel()
        h_tv = torch.pow(x[:, :, 1:, :] - x[:, :, :-1, :], 2).sum()
        w_tv = torch.pow(x[:, :, :, 1:] - x[:, :, :, :-1], 2).sum()
        return self.TVLoss_weight * 2 * (h_tv / n_h + w_tv / n_w) / x.shape[0]


def oracle_spec():
    """FieldSpec of the same workload for the CPU legs (the only place bench.py touches oracle/)."""
    from oracle import t2n_oracle as orc
    return orc.FieldSpec(aabb=AABB, grid=GRID, near_far=NEAR_FAR, step_ratio=STEP_RATIO)


def view_pose(rank):
    """Camera at the origin looking down +z, yawed a little per rank so ranks render different views."""
    import math
    a = 0.04 * rank
    return torch.tensor([[math.cos(a), 0.0, math.sin(a), 0.0],
                         [0.0, 1.0, 0.0, 0.0],
                         [-math.sin(a), 0.0, math.cos(a), 0.0]], dtype=torch.float32)


def pinhole_rays(h, w, focal, c2w):
    """[h*w, 6] rays of a pin-hole view on the host: pixel centres at +0.5, OpenCV axes, normalised directions
    (dataLoader/ray_utils.py:24-42, scene_gen.py:45), rotated by the pose without renormalisation (ray_utils.py:66-87)."""
    ys, xs = torch.meshgrid(torch.linspace(0, h - 1, h), torch.linspace(0, w - 1, w), indexing="ij")
    d = torch.stack([(xs + 0.5 - w / 2) / focal, (ys + 0.5 - h / 2) / focal, torch.ones_like(xs)], -1)
    d = d / torch.norm(d, dim=-1, keepdim=True)
    rd = d @ c2w[:3, :3].T
    ro = c2w[:3, 3].expand(rd.shape)
    return torch.cat([ro.reshape(-1, 3), rd.reshape(-1, 3)], -1).contiguous()


def host_rays(rank):
    return pinhole_rays(H, W, FOCAL, view_pose(rank))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_leg(spec, params, rays_cpu, steps, warmup, what):
    """Times the oracle on the host cores: bounded ray sample of the same workload.  Returns the sample size, the step
    times and what the oracle computed on the sample (inputs + outputs), which the product arm compares its own result
    for the SAME rays against (the `parity` block of the JSON line)."""
    from oracle import ref_runner
    from oracle import t2n_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    S = orc.derive_step(spec)[1]
    g = torch.Generator().manual_seed(0)
    n = CPU_RAYS_FWD if what == "fwd" else CPU_RAYS_BWD
    idx = torch.randperm(rays_cpu.shape[0], generator=g)[:n]
    rays = rays_cpu[idx].contiguous()
    times = []
    # oracle/_ref (the unmodified reference modules, oracle/make_ref.py) when it travelled with the snapshot, else the port
    ref_model = ref_runner.build(spec, params) if ref_runner.available() else None
    kept = {"rays": rays, "S": S, "kind": "reference" if ref_model is not None else "port"}
    if what == "fwd":
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            with torch.no_grad():
                if ref_model is not None:
                    out = ref_runner.render(ref_model, rays, S, False, True)
                else:
                    out = orc.render(spec, params, rays, S, False, True, None)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        kept.update(rgb_map=out[0], depth_map=out[1], z_vals=out[2], weight=out[3], weight_thres=spec.weight_thres)
    else:
        jitter_seed = 4321
        torch.manual_seed(jitter_seed)
        jitter = torch.rand(n, 1)                        # what tensorBase.py:316 draws after the same seed
        rgb_gt, depth_gt = torch.rand(n, 3, generator=g), 2 + 4 * torch.rand(n, generator=g)
        if ref_model is not None:
            p = dict(ref_model.named_parameters())
        else:
            p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        for i in range(warmup + steps):
            for v in p.values():
                v.grad = None
            t0 = time.perf_counter()
            if ref_model is not None:
                out = ref_runner.render(ref_model, rays, S, True, True, jitter_seed=jitter_seed)
            else:
                out = orc.render(spec, p, rays, S, True, True, jitter)
            loss = orc.training_loss(*out, rgb_gt, depth_gt)
            loss.backward()
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        kept.update(jitter=jitter, rgb_gt=rgb_gt, depth_gt=depth_gt, loss=float(loss.detach()),
                    grads={k: v.grad for k, v in p.items()}, weight=out[3].detach(), weight_thres=spec.weight_thres)
    return n, times, kept


def cpu_reference_regrad(spec, params, kept, selection):
    """Part of the CPU leg (checker role): loss and gradients of the reference's function on a given app-sample
    selection (oracle port, render(app_mask_override=...))."""
    from oracle import t2n_oracle as orc
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    out = orc.render(spec, p, kept["rays"], kept["S"], True, True, kept["jitter"], None, app_mask_override=selection)
    loss = orc.training_loss(*out, kept["rgb_gt"], kept["depth_gt"])
    loss.backward()
    return dict(kept, loss=float(loss.detach()), grads={k: v.grad for k, v in p.items()})


def cpu_reference_kinks(spec, params, kept):
    """Part of the CPU leg (checker role): listed samples of the training sample that sit within 4e-6 of a ReLU kink of
    the decoder (oracle.relu_kink_samples) -- there the derivative may legitimately be taken on either side."""
    from oracle import t2n_oracle as orc
    with torch.no_grad():
        aux = orc.render(spec, params, kept["rays"], kept["S"], True, True, kept["jitter"], None, keep=True)[4]
    return orc.relu_kink_samples(spec, params, kept["rays"], aux)[0]


def _psnr(a, b):
    import math
    mse = float(((a.double() - b.double()) ** 2).mean())
    return 200.0 if mse == 0 else -10.0 * math.log10(mse)


def parity_forward(model, dev, kept):
    """The product arm's render of the oracle's ray sample of the timed view against what the oracle computed:
    models/tensorBase.py:436-507 on identical rays and field.  Gates (BASELINE.json north_star): RGB 1e-4 relative,
    PSNR within 0.05 dB; z_vals bit-exact."""
    with torch.no_grad():
        rgb, depth, z, w = [t.cpu() for t in model(kept["rays"].to(dev), is_train=False, white_bg=True, ndc_ray=0,
                                                   N_samples=kept["S"])]
    ref_rgb, ref_w = kept["rgb_map"], kept["weight"]
    g = torch.Generator().manual_seed(5)
    target = torch.rand(ref_rgb.shape, generator=g)          # any fixed image: PSNR(ours, t) against PSNR(reference, t)
    out = {"sample": f"{ref_rgb.shape[0]} rays of the timed 800x800 view, S={kept['S']}, eval forward (the cpu_baseline sample)",
           "rgb_max_rel": float(((rgb - ref_rgb).abs() / ref_rgb.abs().clamp_min(0.05)).max()),
           "rgb_max_abs": float((rgb - ref_rgb).abs().max()),
           "depth_max_rel": float(((depth - kept["depth_map"]).abs() / kept["depth_map"].abs().clamp_min(0.05)).max()),
           "weight_max_abs": float((w - ref_w).abs().max()),
           "z_bit_exact": bool(torch.equal(z, kept["z_vals"])),
           "app_mask_flips": int(((w > kept["weight_thres"]) != (ref_w > kept["weight_thres"])).sum()),
           "listed_samples": int((ref_w > kept["weight_thres"]).sum()),
           "psnr_db": _psnr(rgb, ref_rgb),
           "psnr_delta_db": abs(_psnr(rgb, target) - _psnr(ref_rgb, target))}
    out["ok"] = bool(out["rgb_max_rel"] <= 1e-4 and out["z_bit_exact"] and out["psnr_delta_db"] <= 0.05
                     and out["weight_max_abs"] <= 2e-6 and out["depth_max_rel"] <= 1e-4)
    return out


def parity_backward(model, dev, kept, spec, params):
    """Loss and all parameter gradients of the Text2NeRF data loss (text2nerf_main.py:563-575) on the oracle's training
    sample, through the fused data_loss path with the oracle's jitter.  The selection weight > rayMarch_weight_thres
    (tensorBase.py:477) is a discontinuity of the reference's function: when isolated samples within fp32 noise of the
    threshold are selected differently, the reference gradients are re-evaluated (oracle port) on the product arm's
    selection, and the number of such samples is reported."""
    from text2nerf_b200.tensorBase import _FusedLossFn, _RenderFn
    R, S = kept["rays"].shape[0], kept["S"]
    with torch.no_grad():
        w = _RenderFn.apply(model, kept["rays"].to(dev), kept["jitter"].reshape(-1).to(dev), S, True, True, False,
                            *model._flat_params())[3].cpu()
    sel = w > kept["weight_thres"]
    flips = int((sel != (kept["weight"] > kept["weight_thres"])).sum())
    if 0 < flips <= 8:
        kept = cpu_reference_regrad(spec, params, kept, sel)
    kinks = cpu_reference_kinks(spec, params, kept)
    model.zero_grad()
    worst, worst_cos, worst_l2, per = 0.0, 1.0, 0.0, {}
    loss = None
    for _ in range(2):      # the first pass sizes the tensor-core backward's capacity; the second is the one compared
        model.zero_grad()
        loss = _FusedLossFn.apply(model, kept["rays"].to(dev), kept["jitter"].reshape(-1).to(dev), S, True,
                                  kept["rgb_gt"].to(dev), kept["depth_gt"].to(dev), 0.005, 1e3, 0.1, 1.0 / R,
                                  True, *model._flat_params())[0]
        loss.backward()
        torch.cuda.synchronize()
    for k, p in model.named_parameters():
        a, b = p.grad.detach().double().cpu().flatten(), kept["grads"][k].double().flatten()
        err = float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
        cos = float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-300))
        per[k] = [err, cos]
        worst, worst_cos = max(worst, err), min(worst_cos, cos)
        worst_l2 = max(worst_l2, float((a - b).norm() / b.norm().clamp_min(1e-300)))
    model.zero_grad()
    out = {"sample": f"{R} rays, S={S}, fused data_loss + backward, {len(per)} parameter tensors",
           "loss_rel": abs(float(loss) - kept["loss"]) / abs(kept["loss"]),
           "grad_max_scaled_err": worst, "grad_min_cosine": worst_cos, "grad_max_rel_l2": worst_l2, "n_grads": len(per),
           "app_mask_flips": flips, "relu_kink_samples": kinks,
           "gates": "loss 2e-5; no sample on a ReLU kink (|h| < 4e-6): max err <= 2e-4 of scale, cosine > 1-1e-6; "
                    "otherwise (the derivative of that unit may be taken on either side): cosine > 1-1e-5, rel L2 <= 5e-3"}
    if kinks == 0:
        grads_ok = worst <= 2e-4 and worst_cos > 1 - 1e-6
    else:
        grads_ok = worst_cos > 1 - 1e-5 and worst_l2 <= 5e-3
    out["ok"] = bool(out["loss_rel"] <= 2e-5 and grads_ok)
    return out


def run_reference_arm(args, rank):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores: the unmodified
    reference modules under oracle/_ref (oracle/make_ref.py; kind "reference"), else the oracle port (same ATen ops, tests
    pin it bit-exact to the unmodified reference; kind "port")."""
    if rank != 0:
        return
    spec = oracle_spec()
    params = make_params()
    rays = host_rays(0)
    n, times, kept = cpu_reference_leg(spec, params, rays, args.steps, max(1, min(args.warmup, 1)), "fwd")
    total = sum(times)
    v = n * len(times) / total / 1e6
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "Mrays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "lego800: 300^3 VM field, 800x800 view, S=1036 (bounded CPU sample per step)"},
            "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": kept["kind"],
                             "sample": f"{n} random rays of the 800x800 view per step, S=1036, eval forward"},
            "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def shard_equivalence(model, flat, dev, rank, world, S):
    """max over the 19 gradient tensors of |all-reduced sharded gradient - single-rank full-batch gradient| / max |.| on one
    4096-ray batch of view 0 (identical rays, jitter and targets on every rank; each rank differentiates its slice)."""
    import torch.distributed as dist
    from text2nerf_b200 import dist as t2n_dist
    from text2nerf_b200.tensorBase import _FusedLossFn
    g = torch.Generator().manual_seed(4242)
    R = TRAIN_BATCH
    full = host_rays(0)
    rays = full[torch.randint(0, full.shape[0], (R,), generator=g)].contiguous().to(dev)
    jitter = torch.rand(R, generator=g).to(dev)
    rgb_gt, depth_gt = torch.rand(R, 3, generator=g).to(dev), (2 + 4 * torch.rand(R, generator=g)).to(dev)

    def run(lo, hi):
        flat.zero_()
        loss = _FusedLossFn.apply(model, rays[lo:hi].contiguous(), jitter[lo:hi].contiguous(), S, True,
                                  rgb_gt[lo:hi].contiguous(), depth_gt[lo:hi].contiguous(), 0.005, 1e3, 0.1, 1.0 / R, True,
                                  *model._flat_params())[0]
        loss.backward()

    lo, hi = t2n_dist.shard_bounds(R, rank, world)
    for _ in range(2):          # first pass sizes the tensor-core backward's operand images
        run(lo, hi)
    t2n_dist.allreduce_flat_grads_overlapped(model, world)
    torch.cuda.synchronize()
    sharded = flat.clone()
    err = torch.zeros(1, device=dev)
    if rank == 0:
        for _ in range(2):
            run(0, R)
        torch.cuda.synchronize()
        worst_t = 0.0
        for v in model._flat_grad["views"]:
            off = (v.data_ptr() - flat.data_ptr()) // 4
            a, b = sharded[off:off + v.numel()], flat[off:off + v.numel()]
            worst_t = max(worst_t, float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)))
        err[0] = worst_t
    dist.broadcast(err, 0)
    return float(err)


def strong_scaling_leg(params, dev, rank, world, timed, steps):
    """Text2NeRF's own training shape (aabb +-8, 300^3, step_ratio 1.0 -> S = 259, ONE 16384-ray batch per step;
    text2nerf_main.py:437-439, 662-664) with the batch split over the ranks: strong scaling of a fixed batch."""
    import contextlib
    import io
    from text2nerf_b200 import TensorVMSplit
    from text2nerf_b200 import dist as t2n_dist
    R = 16384
    with contextlib.redirect_stdout(io.StringIO()):
        m = TensorVMSplit(torch.tensor([[-8.0, -8, -8], [8, 8, 8]], device=dev), GRID, dev, density_n_comp=[16, 16, 16],
                          appearance_n_comp=[48, 48, 48], app_dim=27, near_far=[0.5, 8.0], shadingMode="MLP_Fea_noview",
                          alphaMask_thres=0.001, density_shift=-10, distance_scale=25, pos_pe=6, view_pe=2, fea_pe=6,
                          featureC=128, step_ratio=1.0, fea2denseAct="softplus")
    m.load_state_dict({k: v.to(dev) for k, v in params.items()})
    S = m.nSamples // 2
    flat = m.enable_flat_grads(True)
    if world > 1:
        t2n_dist.enable_overlapped_allreduce(m)
    g = torch.Generator().manual_seed(77)
    pose = torch.tensor([[1.0, 0, 0, 0.1], [0, 1.0, 0, -0.05], [0, 0, 1.0, 0.2]])
    view = pinhole_rays(512, 512, 512.0, pose)
    lo, hi = t2n_dist.shard_bounds(R, rank, world)
    batches = []
    for _ in range(4):
        idx = torch.randint(0, view.shape[0], (R,), generator=g)
        rgb, dep = torch.rand(R, 3, generator=g), 0.5 + 7.5 * torch.rand(R, generator=g)
        batches.append((view[idx][lo:hi].contiguous().to(dev), rgb[lo:hi].contiguous().to(dev), dep[lo:hi].contiguous().to(dev)))
    torch.manual_seed(11)

    def step():
        for rays_b, rgb_gt, depth_gt in batches:
            flat.zero_()
            m.data_loss(rays_b, rgb_gt, depth_gt, white_bg=True, N_samples=S, n_rays_total=R).backward()
            t2n_dist.allreduce_flat_grads_overlapped(m, world)

    for _ in range(3):
        step()
    ms = timed(step, steps)
    per_batch = ms / (steps * len(batches))
    n_app, n_valid = m.app_sample_count()
    out = {"what": "ONE 16384-ray Text2NeRF training batch (aabb +-8, 300^3, S=259) per step, rays split over the ranks, "
                   "fused data loss fwd+bwd + the two-segment gradient all-reduce",
           "total_rays_per_batch": R, "rays_per_rank": hi - lo, "ms_per_batch": per_batch, "value": R / per_batch / 1e3,
           "unit": "Mrays/s", "scaling": "strong", "listed_per_ray": n_app / max(1, hi - lo), "valid_per_ray": n_valid / max(1, hi - lo)}
    m.enable_flat_grads(False)
    del m
    torch.cuda.empty_cache()
    return out


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    # keep stdout to the single JSON line: NCCL prints its version banner there at NCCL_DEBUG=VERSION
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    import torch.distributed as dist
    from text2nerf_b200 import OctreeRender_trilinear_fast, TensorVMSplit, _native as nat
    from text2nerf_b200 import dist as t2n_dist
    from text2nerf_b200 import ray_utils

    assert torch.cuda.is_available(), "bench.py (impl ours) needs a CUDA device; there is no CPU fallback"
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nat.load()

    params = make_params()
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        model = TensorVMSplit(torch.tensor(AABB, dtype=torch.float32, device=dev), GRID, dev, density_n_comp=[16, 16, 16],
                              appearance_n_comp=[48, 48, 48], app_dim=27, near_far=NEAR_FAR,
                              shadingMode="MLP_Fea_noview", alphaMask_thres=0.001, density_shift=-10,
                              distance_scale=25, pos_pe=6, view_pe=2, fea_pe=6, featureC=128,
                              step_ratio=STEP_RATIO, fea2denseAct="softplus")
    model.load_state_dict({k: v.to(dev) for k, v in params.items()})
    S = model.nSamples
    assert S == 1036

    pose = view_pose(rank)
    rays_dev = ray_utils.camera_rays(pose, H, W, [FOCAL, FOCAL], normalize=True, device=dev)
    rays_pinned = host_rays(rank).pin_memory()
    n_rays = H * W

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    # ---------------- forward, rays resident in HBM
    def fwd_resident():
        with torch.no_grad():
            model(rays_dev, is_train=False, white_bg=True, ndc_ray=0, N_samples=S)

    # ---------------- forward, end to end through the reference-facing API
    host_rgb = torch.empty((n_rays, 3), dtype=torch.float32).pin_memory()
    host_depth = torch.empty((n_rays,), dtype=torch.float32).pin_memory()

    def fwd_e2e():
        with torch.no_grad():
            rgb, _, depth, _, _ = OctreeRender_trilinear_fast(rays_pinned, model, chunk=n_rays, N_samples=S,
                                                              ndc_ray=0, white_bg=True, is_train=False, device=dev)
        host_rgb.copy_(rgb, non_blocking=True)
        host_depth.copy_(depth, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(max(args.warmup, 3)):
        fwd_resident()
    with ClockSampler(local_rank) as clk:
        ms_fwd = timed(fwd_resident, args.steps)
    value = world * n_rays * args.steps / (ms_fwd * 1e-3) / 1e6

    for _ in range(2):
        fwd_e2e()
    ms_e2e = timed(fwd_e2e, args.steps)
    e2e_value = world * n_rays * args.steps / (ms_e2e * 1e-3) / 1e6

    # ---------------- per-kernel roofline (profiled steps outside the timed region)
    lib = nat.load()
    lib.t2n_profile_enable(1)
    ktimes = {}
    for _ in range(3):
        fwd_resident()
        for name, ms in nat.profile_read():
            ktimes.setdefault(name, []).append(ms)
    lib.t2n_profile_enable(0)
    n_app, n_valid = model.app_sample_count()
    kbytes = {"march": n_rays * (24 + 8 * S) + n_valid * 1152,
              "appearance": n_app * (3456 + 4 + 4 + 12),
              "finalize": n_rays * (8 + 8 + 16) + n_app * (4 + 4 + 12)}
    kavg = {k: sum(v) / len(v) for k, v in ktimes.items()}
    dom = max(kavg, key=kavg.get)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    # DRAM traffic of the dominant kernel per launch: dram__bytes_read.sum + dram__bytes_write.sum of the committed
    # `ncu --set full` capture of this same workload (bench.py never runs under the profiler)
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        if dom in tj:
            traffic, traffic_src = float(tj[dom]["dram_bytes_per_launch"]), tj[dom]["source"]
    except (OSError, ValueError, KeyError):
        pass
    achieved = kbytes[dom] / (kavg[dom] * 1e-3) / 1e9
    total_bytes = n_rays * (40 + 8 * S) + n_valid * 1152 + n_app * 3456
    sm_hz = 1e6 * float((clk.summary().get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0))
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    ncu = {}
    try:
        ncu = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except (OSError, ValueError):
        pass

    def kernel_line(name, ms, alg_bytes, compulsory_bytes, bound, mma_cycles_per_sm=None):
        """One kernel against the roofs that can bind it.  `algorithmic_*`: SURVEY.md 8d bytes (every gathered texel counted
        as if it came from HBM) / CUDA-event time -- it may exceed the HBM peak because the 69 MB of factors are L2/L1
        resident, so it is NOT called a roofline fraction unless HBM is what binds the kernel.  `compulsory_hbm_frac`:
        bytes that must cross HBM (ray inputs, [R,S] outputs, every touched factor once) / time / peak.
        `tensor_pipe_frac`: cycles the issued tcgen05.mma instructions occupy the tensor pipe (3xTF32: 3 MMAs per K step of
        8; 64 cycles at N = 128, 48 at N = 32, tools/mma_rate.cu) / elapsed SM cycles.  ncu_*: the committed
        `ncu --set full` capture of this workload (profiles/ncu_traffic.json)."""
        d = {"ms": ms, "bound": bound, "algorithmic_GBps": alg_bytes / (ms * 1e-3) / 1e9,
             "algorithmic_over_hbm_peak": alg_bytes / (ms * 1e-3) / 1e9 / peak,
             "compulsory_hbm_frac": compulsory_bytes / (ms * 1e-3) / 1e9 / peak}
        if mma_cycles_per_sm is not None:
            d["tensor_pipe_frac"] = mma_cycles_per_sm / (ms * 1e-3 * sm_hz)
        for k, v in ncu.get(name, {}).items():
            if k != "source":
                d["ncu_" + k] = v
        return d

    factor_bytes = 4 * sum(v.numel() for k, v in params.items() if "plane" in k or "line" in k)
    tiles_per_sm = n_app / 128.0 / sms
    # per 128-sample tile: basis 5 chunks at N = 32, layer 1 13 chunks + layer 2 4 chunks at N = 128; 12 MMAs per chunk
    mma_fwd = tiles_per_sm * 12 * (5 * 48 + 17 * 64)
    kernels = {
        "march": kernel_line("march", kavg["march"], kbytes["march"], n_rays * (24 + 8 * S) + factor_bytes * 0.25,
                             "L1 data-pipe wavefronts of the bilinear gathers + issue (factors are L2/L1 resident; HBM only "
                             "carries the [R,S] outputs)"),
        "appearance": kernel_line("appearance", kavg["appearance"], kbytes["appearance"], n_app * 20 + factor_bytes * 0.75,
                                  "decoder warps (8 warps produce the 17 TMEM A chunks of a tile at ~880 cycles per chunk, of which ~450 "
                                  "are the latencies of the hand-off instructions, against 768 cycles of MMAs: tools/trace_mma2.py); "
                                  "tensor pipe and HBM both have headroom", mma_fwd),
        "finalize": kernel_line("finalize", kavg["finalize"], kbytes["finalize"], kbytes["finalize"], "hbm"),
    }
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                "note": "frac = algorithmic gather bytes of the dominant kernel (SURVEY.md 8d: 3476 B per listed sample) / its "
                        "CUDA-event time / measured HBM copy peak, as the contract defines it; what actually binds each kernel is "
                        "in `kernels[*].bound` with the matching utilisation",
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                "kernel_ms": kavg, "kernel_algorithmic_bytes": kbytes, "kernels": kernels,
                "path_algorithmic_GBps": total_bytes / (ms_fwd / args.steps * 1e-3) / 1e9,
                "path_algorithmic_over_hbm_peak": total_bytes / (ms_fwd / args.steps * 1e-3) / 1e9 / peak,
                "valid_samples_per_ray": n_valid / n_rays, "app_samples_per_ray": n_app / n_rays}

    # ---------------- forward + backward on training batches
    fwd_bwd = None
    launches_train = 0
    if not args.no_train:
        g = torch.Generator().manual_seed(100 + rank)
        flat = model.enable_flat_grads(True)
        batches = []
        for _ in range(TRAIN_BATCHES_PER_STEP):
            idx = torch.randint(0, n_rays, (TRAIN_BATCH,), generator=g)
            batches.append((rays_dev[idx.to(dev)].contiguous(), torch.rand(TRAIN_BATCH, 3, generator=g).to(dev),
                            (2 + 4 * torch.rand(TRAIN_BATCH, generator=g)).to(dev)))
        torch.manual_seed(7 + rank)
        if world > 1:
            t2n_dist.enable_overlapped_allreduce(model)
        n_total = world * TRAIN_BATCH       # the means of the loss run over the rays of ALL ranks: no 1/world pass afterwards

        def train_step():           # fused data loss (TensorBase.data_loss): forward, loss kernel, backward kernels
            for rays_b, rgb_gt, depth_gt in batches:
                flat.zero_()
                model.data_loss(rays_b, rgb_gt, depth_gt, white_bg=True, N_samples=S, n_rays_total=n_total).backward()
                t2n_dist.allreduce_flat_grads_overlapped(model, world)

        def train_step_composed():  # the same step written as the reference writes it: tensor ops on the four outputs
            for rays_b, rgb_gt, depth_gt in batches:
                flat.zero_()
                out = model(rays_b, is_train=True, white_bg=True, ndc_ray=0, N_samples=S)
                composed_loss(*out, rgb_gt, depth_gt).backward()
                t2n_dist.allreduce_flat_grads(model, world)

        for _ in range(max(args.warmup, 3)):
            train_step()
        ms_tr = timed(train_step, args.steps)
        for _ in range(2):
            train_step_composed()
        ms_tr_c = timed(train_step_composed, args.steps)
        rays_done = world * TRAIN_BATCH * TRAIN_BATCHES_PER_STEP * args.steps
        fwd_bwd = {"value": rays_done / (ms_tr * 1e-3) / 1e6, "unit": "Mrays/s", "batch_rays_per_gpu": TRAIN_BATCH,
                   "ms_per_batch": ms_tr / (args.steps * TRAIN_BATCHES_PER_STEP),
                   "loss": "rgb MSE + 0.005 depth MSE + 1e3 transmittance (text2nerf_main.py:563-575) through the fused "
                           "TensorBase.data_loss, no optimiser step",
                   "composed_autograd_value": rays_done / (ms_tr_c * 1e-3) / 1e6,
                   "collective": ("all-reduce(sum) of the flat fp32 gradient buffer per batch in two segments: appearance factors (75 %) "
                                  "on a communication stream as soon as T2NGrads::app_done_event fires (behind the scatter kernel), "
                                  "overlapping the weight-gradient GEMMs and the density sweep; density factors + basis + decoder "
                                  "after the backward joined; 1/world folded into the loss")
                   if world > 1 else "none (1 GPU)"}
        lib.t2n_profile_enable(1)
        rays_b, rgb_gt, depth_gt = batches[0]
        flat.zero_()
        loss = model.data_loss(rays_b, rgb_gt, depth_gt, white_bg=True, N_samples=S)
        fk = dict(nat.profile_read())
        loss.backward()
        bk = dict(nat.profile_read())
        lib.t2n_profile_enable(0)
        fwd_bwd["kernel_ms"] = {**fk, **bk}
        fwd_bwd["kernel_ms_note"] = ("profiled batch runs the backward on ONE stream (clean per-kernel event times); the timed "
                                     "region forks the ray sweep and the scatter/dBasis branch onto side streams")
        # roofline of the training batch: algorithmic bytes of SURVEY.md 8d (forward + the same gather bytes again for the
        # scatter-add + the [R,S] state the backward reads) / time of the whole fwd+bwd / HBM peak
        b_app, b_valid = model.app_sample_count()
        tb = TRAIN_BATCH * 16 * S + b_valid * 2 * 1152 + b_app * 2 * 3476
        ms_b = ms_tr / (args.steps * TRAIN_BATCHES_PER_STEP)
        km = fwd_bwd["kernel_ms"]
        fwd_bwd["roofline"] = {
            "bound": "hbm (denominator); the batch is bound by the L2 atomic / scatter rate and producer latency, see kernels",
            "algorithmic_bytes_per_batch": tb, "achieved": tb / (ms_b * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
            "frac": tb / (ms_b * 1e-3) / 1e9 / peak, "valid_samples": b_valid, "listed_samples": b_app,
            "kernels": {k: {"ms": v, "algorithmic_over_hbm_peak": bts / (v * 1e-3) / 1e9 / peak} for k, v, bts in (
                ("march", km.get("march", 0.0), TRAIN_BATCH * (24 + 16 * S) + b_valid * 1152),
                ("appearance", km.get("appearance", 0.0), b_app * 3476),
                ("app_backward_mma", km.get("app_backward_mma", 0.0), b_app * 2.8e3),
                ("app_scatter", km.get("app_scatter", 0.0), b_app * 2 * 3476),
                ("wgrad", km.get("wgrad", 0.0), b_app * 6.4e3),
                ("ray_backward", km.get("ray_backward", 0.0), TRAIN_BATCH * 16 * S + b_valid * 2 * 1152)) if v > 0}}
        # ---- the whole training iteration of text2nerf_main.py:547-598: data loss + TV regularisers + Adam
        # (TV_weight_density 0.1, TV_weight_app 0.01, configs/text2nerf_scenes.txt:31-32), fused against composed
        from text2nerf_b200.optim import FusedAdam

        tvreg = TVLoss()

        def make_iteration(fused):
            opt = (FusedAdam if fused else torch.optim.Adam)(model.get_optparam_groups(0.02, 0.001), betas=(0.9, 0.99))
            reg = tvreg if fused else tvreg.forward      # a plain callable takes the tensor-op route

            def iteration():
                for rays_b, rgb_gt, depth_gt in batches:
                    flat.zero_()
                    if fused:
                        # the parameter-only TV terms are identical on every rank: scaled by 1/world they survive the
                        # summing all-reduce unchanged
                        total = model.data_loss(rays_b, rgb_gt, depth_gt, white_bg=True, N_samples=S, n_rays_total=n_total)
                        total = total + (model.TV_loss_density(reg) * 0.1 + model.TV_loss_app(reg) * 0.01) / world
                        total.backward()
                        t2n_dist.allreduce_flat_grads_overlapped(model, world)
                    else:
                        total = composed_loss(*model(rays_b, is_train=True, white_bg=True, ndc_ray=0, N_samples=S),
                                              rgb_gt, depth_gt)
                        total = total + model.TV_loss_density(reg) * 0.1 + model.TV_loss_app(reg) * 0.01
                        total.backward()
                        t2n_dist.allreduce_flat_grads(model, world)
                    t2n_dist.attach_flat_grads(model)
                    opt.step()
            return iteration

        state0 = {k: v.detach().clone(memory_format=torch.preserve_format) for k, v in model.state_dict().items()}
        full = {}
        for tag, fused in (("fused", True), ("composed", False)):
            it_fn = make_iteration(fused)
            it_fn()
            ms_it = timed(it_fn, 2)
            full[tag + "_ms_per_iteration"] = ms_it / (2 * TRAIN_BATCHES_PER_STEP)
            for p_ in model.parameters():
                p_.grad = None
            model.load_state_dict(state0)
        full["what"] = ("one iteration of text2nerf_main.py:547-598 on a 4096-ray batch: data loss fwd+bwd, TV_loss_density*0.1 + "
                        "TV_loss_app*0.01, Adam step; fused = data_loss + t2n_tv_* + FusedAdam, composed = tensor-op loss and "
                        "TV on the same render kernels + torch.optim.Adam")
        fwd_bwd["full_iteration"] = full
        launches_train = 17 * TRAIN_BATCHES_PER_STEP * args.steps    # march, pack, app, finalize, data_loss | pack_bwd, bwd-data, app_scatter, 4 wgrad, (pack_w1, ffma fallback, unpack: early exit), ray_backward
        # ---- N > 1: the sharded step must equal the single-GPU step on the full batch
        if world > 1:
            fwd_bwd["shard_equivalence_err"] = shard_equivalence(model, flat, dev, rank, world, S)
        # ---- strong scaling: ONE 16384-ray Text2NeRF training batch (text2nerf_main.py:662-664) split over the ranks
        fwd_bwd["strong"] = strong_scaling_leg(params, dev, rank, world, timed, args.steps)
        if world > 1:
            t2n_dist.enable_overlapped_allreduce(model, False)
        model.enable_flat_grads(False)

    # ---------------- CPU baseline (rank 0, N=1 only) + parity of the product arm on the baseline's own ray samples
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        spec = oracle_spec()
        n, times, kept = cpu_reference_leg(spec, params, rays_pinned, 2, 1, "fwd")
        cpu = {"value": n / min(times) / 1e6, "unit": "Mrays/s", "cores": torch.get_num_threads(), "kind": kept["kind"],
               "sample": f"{n} random rays of the same 800x800 view, S=1036, eval forward, best of 2 after 1 warm-up"}
        parity = {"forward": parity_forward(model, dev, kept)}
        if not args.no_train:
            nb, tb, kept_b = cpu_reference_leg(spec, params, rays_pinned, 1, 1, "bwd")
            cpu["fwd_bwd_value"] = nb / min(tb) / 1e6
            cpu["fwd_bwd_sample"] = f"{nb} rays, S=1036, loss of text2nerf_main.py:563-575 + backward"
            parity["fwd_bwd"] = parity_backward(model, dev, kept_b, spec, params)
        parity["ok"] = all(v["ok"] for v in parity.values())

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_fwd / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "lego800: TensorVMSplit 300^3 (16/48 comps, MLP_Fea_noview 351-128-128-3), "
                                       "aabb +-1.5 at z 2.5..5.5, 800x800 view = 640000 rays/step/GPU, S=1036, eval",
                           "rays_per_step_per_gpu": n_rays, "samples_per_ray": S, "parallelism": f"view-sharded x{world}",
                           "l2": "per-step working set ~16 GB (outputs [R,S] stream through) >> 126 MB L2; no explicit flush; "
                                 "the 69 MB of VM factors are L2-resident by design"},
                "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": n_rays * 24,
                        "d2h_bytes_per_step": n_rays * 16, "ms_per_step": ms_e2e / args.steps,
                        "api": "text2nerf_b200.OctreeRender_trilinear_fast(pinned host rays) + rgb/depth .cpu()"},
                "gpu_launches": 4 * args.steps + launches_train,
                "clocks": clk.summary(), "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "fwd_bwd": fwd_bwd}
        print(json.dumps(line))
        if parity is not None and not parity["ok"]:
            print("bench.py: PARITY FAILED against the oracle on the timed workload: " + json.dumps(parity), file=sys.stderr)
            sys.exit(3)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
