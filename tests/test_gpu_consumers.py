"""Render-side consumers against goldens the UNMODIFIED reference produced on CPU
(tests/golden/make_golden_f4.py): Warper.forward_warp (scripts/Warper.py:21-172, numpy float64),
sparse_bilateral_filtering (dataLoader/bilateral_filtering.py:5-35, 138-186) and the per-view assembly of
renderer.evaluation (renderer.py:85-101, 112)."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN_DIR

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def f4():
    return np.load(os.path.join(GOLDEN_DIR, "f4_consumers.npz"))


@pytest.mark.parametrize("tag", ["a", "b"])
def test_forward_warp_vs_reference(tag, f4, cuda_device):
    from text2nerf_b200.consumers import Warper
    mask = f4[f"warp_{tag}/mask"] if f"warp_{tag}/mask" in f4.files else None
    frame2, mask2, depth2, flow = Warper(device=cuda_device).forward_warp(
        f4[f"warp_{tag}/frame"], mask, f4[f"warp_{tag}/depth"], f4[f"warp_{tag}/T1"], f4[f"warp_{tag}/T2"], f4[f"warp_{tag}/K"], None)
    ref_f, ref_m, ref_d, ref_flow = (f4[f"warp_{tag}/out_frame"], f4[f"warp_{tag}/out_mask"], f4[f"warp_{tag}/out_depth"],
                                     f4[f"warp_{tag}/flow"])
    assert np.abs(flow - ref_flow).max() <= 1e-9
    assert np.array_equal(mask2, ref_m)                        # byte work: the known-pixel mask is exact
    assert np.abs(depth2 - ref_d).max() <= 1e-9 * max(1.0, np.abs(ref_d).max())
    # uint8 image: float64 sums in another order may move a value across a .5 rounding boundary -- at most a handful
    diff = np.abs(frame2.astype(np.int32) - ref_f.astype(np.int32))
    assert diff.max() <= 1 and int((diff > 0).sum()) <= 3, (int(diff.max()), int((diff > 0).sum()))
    # tensor in -> tensor out, same values
    out_t = Warper().forward_warp(torch.from_numpy(f4[f"warp_{tag}/frame"]).to(cuda_device), None if mask is None else
                                  torch.from_numpy(mask).to(cuda_device), torch.from_numpy(f4[f"warp_{tag}/depth"]).to(cuda_device),
                                  f4[f"warp_{tag}/T1"], f4[f"warp_{tag}/T2"], f4[f"warp_{tag}/K"])
    assert out_t[0].is_cuda and np.array_equal(out_t[1].cpu().numpy(), ref_m)


def test_sparse_bilateral_filtering_vs_reference(f4, cuda_device):
    from text2nerf_b200.consumers import sparse_bilateral_filtering
    imgs, deps = sparse_bilateral_filtering(f4["bf/depth"].copy(), f4["bf/image"].copy(), filter_size=[7, 5, 5, 3, 3],
                                            depth_threshold=0.02, num_iter=5, HR=False, mask=None, device=cuda_device)
    assert len(imgs) == 5 and len(deps) == 5
    for i in range(5):
        # a weighted median SELECTS input values: bit-exact
        assert np.array_equal(deps[i], f4[f"bf/depths/{i}"]), i
        assert np.array_equal(imgs[i], f4[f"bf/images/{i}"]), i
    imgs, deps = sparse_bilateral_filtering(f4["bf/depth"].copy(), f4["bf/image"].copy(), filter_size=[7, 7, 5, 5, 5],
                                            depth_threshold=0.04, num_iter=3, HR=False, mask=f4["bfm/mask"], device=cuda_device)
    for i in range(3):
        assert np.array_equal(deps[i], f4[f"bfm/depths/{i}"]), i
        assert np.array_equal(imgs[i], f4[f"bfm/images/{i}"]), i


def test_evaluation_view_assembly(cuda_device):
    """renderer.evaluation's per-view arithmetic restated with numpy on the host (renderer.py:92-101, 112)."""
    from text2nerf_b200.consumers import assemble_view
    g = torch.Generator().manual_seed(0)
    H, W = 37, 53
    rgb = (torch.rand(H * W, 3, generator=g) * 1.2 - 0.1)
    depth = torch.rand(H * W, generator=g) * 6
    gt = torch.rand(H, W, 3, generator=g)
    rgb8, dep, psnr = assemble_view(rgb.to(cuda_device), depth.to(cuda_device), H, W, push_depth=2.0, gt_rgb=gt.to(cuda_device))
    ref_rgb = rgb.clamp(0.0, 1.0).reshape(H, W, 3)
    ref_depth = np.maximum((depth.reshape(H, W) - 2.0 + 0.8).numpy(), 0)
    loss = torch.mean((ref_rgb - gt) ** 2)
    ref_psnr = -10.0 * np.log(loss.item()) / np.log(10.0)
    assert np.array_equal(rgb8.cpu().numpy(), (ref_rgb.numpy() * 255).astype("uint8"))
    assert np.abs(dep.cpu().numpy() - ref_depth).max() <= 1e-6
    assert abs(psnr - ref_psnr) <= 1e-4


def test_evaluation_views_loop(cuda_device):
    from helpers import Case, build_model
    from text2nerf_b200.renderer import evaluation_views
    c = Case("t2n_noview_eval")
    model = build_model(c.spec, c.params, cuda_device)
    rays = c.rays[:64].reshape(1, 64, 6).repeat(3, 1, 1)
    gts = torch.rand(3, 64, 3)
    psnrs, rgbs, depths = evaluation_views(rays, model, (8, 8), all_rgbs=gts, N_vis=-1, N_samples=c.n_samples, device=cuda_device)
    assert len(psnrs) == 3 and rgbs[0].shape == (8, 8, 3) and rgbs[0].dtype == np.uint8 and depths[0].shape == (8, 8)
    ref = c.out["rgb_map"][:64].clamp(0, 1)
    assert np.abs(rgbs[0].reshape(-1, 3).astype(np.float32) / 255 - ref.numpy()).max() <= 1.0 / 255 + 1e-4
