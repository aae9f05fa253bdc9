#!/usr/bin/env python
"""BASELINE.json configs 3 and 5 on one GPU: (a) ray-batch sweep 2^10..2^20 at the lego-shaped 300^3 field, forward
(eval) and forward+backward (fused data loss) Mrays/s with the achieved fraction of the HBM roofline (algorithmic bytes of
SURVEY.md 8d / CUDA-event time / MEASURED_PEAKS hbm_gbs); (b) the T2N training shape (aabb +-8, 300^3, S = 259,
16384-ray batches, camera inside the box).  Prints one JSON object; results are copied into profiles/."""
import contextlib, io, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from text2nerf_b200 import TensorVMSplit, ray_utils
from text2nerf_b200 import _native as nat

dev = torch.device("cuda:0")
peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0)) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def build(aabb, near_far, step_ratio, params):
    with contextlib.redirect_stdout(io.StringIO()):
        m = TensorVMSplit(torch.tensor(aabb, dtype=torch.float32, device=dev), bench.GRID, dev, density_n_comp=[16, 16, 16],
                          appearance_n_comp=[48, 48, 48], app_dim=27, near_far=near_far, shadingMode="MLP_Fea_noview",
                          step_ratio=step_ratio, fea_pe=6, view_pe=2)
    m.load_state_dict({k: v.to(dev) for k, v in params.items()})
    return m


def timeit(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def measure(model, rays, S, reps):
    R = rays.shape[0]
    g = torch.Generator().manual_seed(1)
    rgb_gt, depth_gt = torch.rand(R, 3, generator=g).to(dev), (2 + 4 * torch.rand(R, generator=g)).to(dev)

    def fwd():
        with torch.no_grad():
            model(rays, is_train=False, white_bg=True, N_samples=S)

    def fb():
        for p in model.parameters():
            p.grad = None
        model.data_loss(rays, rgb_gt, depth_gt, white_bg=True, N_samples=S).backward()

    ms_f = timeit(fwd, reps)
    n_app, n_valid = model.app_sample_count()
    ms_fb = timeit(fb, max(2, reps // 2))
    bytes_f = R * (40 + 8 * S) + n_valid * 1152 + n_app * 3456
    kernel_ms = None
    if os.environ.get("SWEEP_KERNELS"):
        lib = nat.load()
        lib.t2n_profile_enable(1)
        for p in model.parameters():
            p.grad = None
        loss = model.data_loss(rays, rgb_gt, depth_gt, white_bg=True, N_samples=S)
        kernel_ms = dict(nat.profile_read())
        loss.backward()
        kernel_ms.update(dict(nat.profile_read()))
        lib.t2n_profile_enable(0)
    return {"kernel_ms": kernel_ms,"rays": R, "S": S, "fwd_ms": ms_f, "fwd_Mrays_s": R / ms_f / 1e3, "fwd_bwd_ms": ms_fb, "fwd_bwd_Mrays_s": R / ms_fb / 1e3,
            "valid_per_ray": n_valid / R, "app_per_ray": n_app / R, "fwd_roofline_frac": bytes_f / (ms_f * 1e-3) / 1e9 / peak}


out = {"peak_GBps": peak}
# (a) sweep at the bench field
model = build(bench.AABB, bench.NEAR_FAR, bench.STEP_RATIO, bench.make_params())
S = model.nSamples
full = ray_utils.camera_rays(bench.view_pose(0), bench.H, bench.W, [bench.FOCAL] * 2, normalize=True, device=dev)
g = torch.Generator().manual_seed(0)
out["sweep_lego_300"] = []
for e in ([] if os.environ.get("SWEEP_T2N_ONLY") else range(10, 21)):
    R = 1 << e
    idx = torch.randint(0, full.shape[0], (R,), generator=g).to(dev)
    rays = full[idx].contiguous()
    out["sweep_lego_300"].append(measure(model, rays, S, 20 if e <= 16 else 4))
del model
torch.cuda.empty_cache()
# (b) T2N training shape: box +-8, step_ratio 1.0 (S = 259 as text2nerf_main.py halves nSamples), camera near the origin
model2 = build([[-8, -8, -8], [8, 8, 8]], [0.5, 8.0], 1.0, bench.make_params())
pose = torch.tensor([[1.0, 0, 0, 0.1], [0, 1.0, 0, -0.05], [0, 0, 1.0, 0.2]])
rays2 = bench.pinhole_rays(512, 512, 512.0, pose).to(dev)
idx = torch.randint(0, rays2.shape[0], (16384,), generator=g).to(dev)
out["t2n_train_shape"] = measure(model2, rays2[idx].contiguous(), model2.nSamples // 2, 10)
out["t2n_view_512"] = measure(model2, rays2.contiguous(), model2.nSamples // 2, 3)
print(json.dumps(out))
