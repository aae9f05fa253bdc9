"""Render-side consumers of a rendered RGB-D view, device-resident mirrors of what the Text2NeRF
loop runs on the CPU between training stages:

  Warper.forward_warp             scripts/Warper.py:21-62     DIBR bilinear splat          -> t2n_forward_warp
  sparse_bilateral_filtering      dataLoader/bilateral_filtering.py:5-35  edge-aware smoothing (weighted median)
                                                                                           -> t2n_depth_discontinuity / t2n_weighted_median
  renderer.evaluation (per view)  renderer.py:85-101,112      clamp, uint8, depth shift, PSNR -> t2n_assemble_view

Same names, argument meaning and return values as the reference functions; inputs may be numpy arrays (as the reference
takes them) or CUDA tensors (a rendered view that never left the device); numpy in -> numpy out, tensor in -> tensor out.
There is no CPU implementation here: without the CUDA library these raise NativeLibraryError."""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import numpy as np
import torch

from . import _native as nat


def _dev(device=None):
    if device is not None:
        return torch.device(device)
    if not torch.cuda.is_available():
        raise nat.NativeLibraryError("text2nerf_b200.consumers needs a CUDA device (no CPU path)")
    return torch.device("cuda", torch.cuda.current_device())


def _to_dev(x, dtype, dev):
    if isinstance(x, torch.Tensor):
        return x.to(dev, dtype).contiguous()
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev, dtype).contiguous()


def _darr(a, n):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64)).reshape(-1)
    assert a.size == n
    return (C.c_double * n)(*a.tolist())


class Warper:
    """scripts/Warper.py: forward_warp only (the other methods are file IO / demo code)."""

    def __init__(self, resolution: tuple = None, device=None):
        self.resolution = resolution
        self.device = device

    def forward_warp(self, frame1, mask1, depth1, transformation1, transformation2, intrinsic1, intrinsic2=None):
        """frame1 (h, w, 3) uint8, mask1 (h, w) bool or None, depth1 (h, w) float, transformation1/2 (4, 4) extrinsics,
        intrinsic1/2 (3, 3).  Returns warped_frame2 (h, w, 3) uint8, mask2 (h, w) bool, warped_depth2 (h, w) float64,
        flow12 (h, w, 2) float64."""
        lib = nat.load()
        as_numpy = not isinstance(frame1, torch.Tensor)
        dev = _dev(self.device if self.device is not None else (None if as_numpy else frame1.device))
        h, w = int(frame1.shape[0]), int(frame1.shape[1])
        if self.resolution is not None:
            assert (h, w) == tuple(self.resolution)
        assert tuple(frame1.shape) == (h, w, 3) and tuple(depth1.shape) == (h, w)
        t1, t2, k1 = np.asarray(transformation1), np.asarray(transformation2), np.asarray(intrinsic1)
        k2 = np.copy(k1) if intrinsic2 is None else np.asarray(intrinsic2)
        assert t1.shape == (4, 4) and t2.shape == (4, 4) and k1.shape == (3, 3) and k2.shape == (3, 3)
        # the two tiny matrix inversions / products of compute_transformed_points stay on the host, in numpy and in the
        # callers' dtypes like the reference (utils.py:90-94 passes a float32 intrinsic matrix: its inverse is rounded to
        # float32 before the float64 per-pixel arithmetic)
        M = np.matmul(t2, np.linalg.inv(t1))
        k1inv = np.linalg.inv(k1)
        frame = _to_dev(frame1, torch.uint8, dev)
        depth = _to_dev(depth1, torch.float64, dev)
        mask = None if mask1 is None else _to_dev(mask1, torch.uint8, dev)
        scratch = torch.empty((int(lib.t2n_forward_warp_scratch_doubles(h, w)),), device=dev, dtype=torch.float64)
        out_frame = torch.empty((h, w, 3), device=dev, dtype=torch.uint8)
        out_mask = torch.empty((h, w), device=dev, dtype=torch.uint8)
        out_depth = torch.empty((h, w), device=dev, dtype=torch.float64)
        flow = torch.empty((h, w, 2), device=dev, dtype=torch.float64)
        with torch.cuda.device(dev):
            rc = lib.t2n_forward_warp(frame.data_ptr(), None if mask is None else mask.data_ptr(), depth.data_ptr(),
                                      _darr(M, 16), _darr(k1inv, 9), _darr(k2, 9), h, w, scratch.data_ptr(),
                                      out_frame.data_ptr(), out_mask.data_ptr(), out_depth.data_ptr(), flow.data_ptr(),
                                      torch.cuda.current_stream(dev).cuda_stream)
        nat.check(rc, "t2n_forward_warp")
        if as_numpy:
            return out_frame.cpu().numpy(), out_mask.cpu().numpy().astype(bool), out_depth.cpu().numpy(), flow.cpu().numpy()
        return out_frame, out_mask.bool(), out_depth, flow


def sparse_bilateral_filtering(depth, image, filter_size=[7, 7, 5, 5, 5], depth_threshold=0.04, num_iter=5, HR=False,
                               mask=None, device=None):
    """dataLoader/bilateral_filtering.py:5-35.  depth (H, W) fp32, image (H, W, 3) fp32.  Returns (save_images,
    save_depths) with the reference's exact semantics, quirk included: save_depths[i] is the depth BEFORE iteration i (the
    reference appends, then rebinds vis_depth to a new array), but every save_images[i] is the SAME array -- the reference
    appends vis_image and then filters it in place (bilateral_filtering.py:16, 30-33) -- i.e. the image after ALL
    num_iter iterations.  Callers take [-1] of both (text2nerf_main.py:115-119)."""
    lib = nat.load()
    as_numpy = not isinstance(depth, torch.Tensor)
    dev = _dev(device if device is not None else (None if as_numpy else depth.device))
    vis_depth = _to_dev(depth, torch.float32, dev).clone()
    depth0 = vis_depth.clone()
    vis_image = _to_dev(image, torch.float32, dev).clone()
    H, W = int(vis_depth.shape[0]), int(vis_depth.shape[1])
    m = None if mask is None else _to_dev(mask, torch.uint8, dev)
    chans = [vis_image[:, :, c].contiguous() for c in range(3)]
    disc = torch.empty((H, W), device=dev, dtype=torch.float32)
    save_images, save_depths = [], []
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        for i in range(num_iter):
            window = int(filter_size[i] if isinstance(filter_size, list) else filter_size)
            save_depths.append(vis_depth)
            nat.check(lib.t2n_depth_discontinuity(vis_depth.data_ptr(), depth0.data_ptr(), None if m is None else m.data_ptr(),
                                                  float(depth_threshold), H, W, disc.data_ptr(), st), "t2n_depth_discontinuity")
            outs = []
            for src in [vis_depth] + chans:
                dst = torch.empty_like(src)
                nat.check(lib.t2n_weighted_median(src.data_ptr(), disc.data_ptr(), None if m is None else m.data_ptr(), H, W,
                                                  window, dst.data_ptr(), st), "t2n_weighted_median")
                outs.append(dst)
            vis_depth, chans = outs[0], outs[1:]
    final_image = torch.stack(chans, -1)
    save_images = [final_image] * num_iter
    if as_numpy:
        img = final_image.cpu().numpy()
        return [img] * num_iter, [t.cpu().numpy() for t in save_depths]
    return save_images, save_depths


def assemble_view(rgb_map, depth_map, H, W, push_depth=2.0, gt_rgb=None):
    """The per-view part of renderer.evaluation (renderer.py:92-96, 98-101, 112) on the device:
    rgb8 (H, W, 3) uint8 = (clamp(rgb_map, 0, 1) * 255).astype(uint8), depth (H, W) = max(depth_map - push_depth + 0.8, 0)
    and, with a ground-truth view, PSNR = -10 ln(mse) / ln 10 of the clamped fp32 image.  Returns (rgb8, depth, psnr)."""
    lib = nat.load()
    if not rgb_map.is_cuda:
        raise nat.NativeLibraryError("assemble_view takes the renderer's CUDA tensors")
    dev = rgb_map.device
    n = H * W
    rgb = rgb_map.detach().reshape(n, 3).float().contiguous()
    dep = depth_map.detach().reshape(n).float().contiguous()
    gt = None if gt_rgb is None else gt_rgb.detach().to(dev, torch.float32).reshape(n, 3).contiguous()
    rgb8 = torch.empty((H, W, 3), device=dev, dtype=torch.uint8)
    depth_out = torch.empty((H, W), device=dev, dtype=torch.float32)
    sq = torch.zeros((1,), device=dev, dtype=torch.float64)
    with torch.cuda.device(dev):
        rc = lib.t2n_assemble_view(rgb.data_ptr(), dep.data_ptr(), None if gt is None else gt.data_ptr(), n,
                                   float(-push_depth + 0.8), rgb8.data_ptr(), depth_out.data_ptr(), sq.data_ptr(),
                                   torch.cuda.current_stream(dev).cuda_stream)
    nat.check(rc, "t2n_assemble_view")
    psnr = None
    if gt is not None:
        mse = float(sq) / (3 * n)
        psnr = -10.0 * math.log(mse) / math.log(10.0) if mse > 0 else float("inf")
    return rgb8, depth_out, psnr
