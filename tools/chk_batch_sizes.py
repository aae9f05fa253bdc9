import sys, os, contextlib, io, torch
sys.path.insert(0, os.getcwd())
import bench
from text2nerf_b200 import TensorVMSplit, ray_utils
dev = torch.device("cuda:0")
params = bench.make_params()
with contextlib.redirect_stdout(io.StringIO()):
    m = TensorVMSplit(torch.tensor(bench.AABB, dtype=torch.float32, device=dev), bench.GRID, dev, density_n_comp=[16,16,16], appearance_n_comp=[48,48,48], app_dim=27, near_far=bench.NEAR_FAR, shadingMode="MLP_Fea_noview", step_ratio=bench.STEP_RATIO, fea_pe=6, view_pe=2)
m.load_state_dict({k: v.to(dev) for k, v in params.items()})
S = m.nSamples
rays = ray_utils.camera_rays(bench.view_pose(0), bench.H, bench.W, [bench.FOCAL]*2, device=dev)
g = torch.Generator().manual_seed(0)
for R in (4096, 8192, 16384, 32768):
    idx = torch.randint(0, rays.shape[0], (R,), generator=g).to(dev)
    rb = rays[idx].contiguous(); rgb = torch.rand(R,3,generator=g).to(dev); dep = (2+4*torch.rand(R,generator=g)).to(dev)
    ts=[]
    for it in range(8):
        m.zero_grad(); torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); m.data_loss(rb, rgb, dep, white_bg=True, N_samples=S).backward(); e1.record(); torch.cuda.synchronize()
        ts.append(round(e0.elapsed_time(e1),2))
    print(R, ts, "Mrays/s best", round(R/min(ts)/1e3,3))
