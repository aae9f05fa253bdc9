"""Host-side mirror of the two ray-generation functions of ``dataLoader/ray_utils.py`` that
feed the hot path: get_ray_directions (:24-42) and get_rays (:66-87).  On a CUDA device both
run as one kernel of libt2n_b200.so; there is no host fallback."""
import ctypes as C

import torch

from . import _native as nat


def _pose12(c2w):
    m = torch.as_tensor(c2w, dtype=torch.float32).detach().cpu()[:3, :4].contiguous()
    return (C.c_float * 12)(*m.view(-1).tolist())


def get_ray_directions(H, W, focal, center=None, device="cuda", normalize=False):
    """(H, W, 3) camera-space directions, pixel centres at +0.5, OpenCV axes.  `focal` is
    [fx, fy] like the reference.  `normalize=True` additionally divides by the norm, which is
    what dataLoader/scene_gen.py:45 does to the result."""
    lib = nat.load()
    fx, fy = float(focal[0]), float(focal[1])
    cx, cy = (W / 2, H / 2) if center is None else (float(center[0]), float(center[1]))
    ident = (C.c_float * 12)(1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0)
    rays = torch.empty((H * W, 6), device=device, dtype=torch.float32)
    with torch.cuda.device(rays.device):
        rc = lib.t2n_get_rays(ident, fx, fy, cx, cy, H, W, int(normalize), rays.data_ptr(),
                              torch.cuda.current_stream(rays.device).cuda_stream)
    nat.check(rc, "t2n_get_rays")
    return rays[:, 3:6].reshape(H, W, 3)


def get_rays(directions, c2w):
    """rays_o, rays_d = (H*W, 3) each: directions rotated by c2w[:3,:3] (no renormalisation),
    origin broadcast from c2w[:3,3]."""
    lib = nat.load()
    d = directions.reshape(-1, 3)
    if not d.is_cuda:
        raise nat.NativeLibraryError("text2nerf_b200.ray_utils.get_rays needs CUDA tensors (no CPU fallback)")
    d = d.float().contiguous()
    rays = torch.empty((d.shape[0], 6), device=d.device, dtype=torch.float32)
    with torch.cuda.device(d.device):
        rc = lib.t2n_rotate_rays(_pose12(c2w), d.data_ptr(), d.shape[0], rays.data_ptr(),
                                 torch.cuda.current_stream(d.device).cuda_stream)
    nat.check(rc, "t2n_rotate_rays")
    return rays[:, :3], rays[:, 3:6]


def camera_rays(c2w, H, W, focal, center=None, normalize=True, device="cuda"):
    """Fused get_ray_directions + normalisation + get_rays + cat: the [H*W, 6] tensor
    renderer.evaluation_path builds per pose (renderer.py:156-163), in one kernel."""
    lib = nat.load()
    fx, fy = float(focal[0]), float(focal[1])
    cx, cy = (W / 2, H / 2) if center is None else (float(center[0]), float(center[1]))
    rays = torch.empty((H * W, 6), device=device, dtype=torch.float32)
    with torch.cuda.device(rays.device):
        rc = lib.t2n_get_rays(_pose12(c2w), fx, fy, cx, cy, H, W, int(normalize), rays.data_ptr(),
                              torch.cuda.current_stream(rays.device).cuda_stream)
    nat.check(rc, "t2n_get_rays")
    return rays
