"""Grid maintenance on the GPU against goldens the UNMODIFIED reference produced on CPU
(tests/golden/make_golden_f3.py): getDenseAlpha / updateAlphaMask (models/tensorBase.py:328-370), filtering_rays in both
modes (:372-404), shrink and upsample_volume_grid (models/tensoRF.py:243-303), run as the sequence a TensoRF-style
coarse-to-fine training performs, plus the packbits checkpoint round trip (tensorBase.py:275-290)."""
import io
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import GOLDEN_DIR, build_model, quiet
from oracle import t2n_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def f3():
    z = np.load(os.path.join(GOLDEN_DIR, "f3_maintenance.npz"))
    d = json.loads(str(z["spec"]))
    d.pop("dtype")
    spec = orc.FieldSpec(**d)
    params = {k[6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param/")}
    return z, spec, params


def _check_mask(model, z, tag, dense_key, thres=0.001):
    """Occupancy volume against the reference's; voxels whose pooled alpha lies within 1e-5 of the threshold may flip
    (the kernel's density differs from the CPU's by ~1e-6 relative)."""
    vol = model.alphaMask.alpha_volume.detach().cpu().view(*z[f"{tag}/volume"].shape)
    ref = torch.from_numpy(z[f"{tag}/volume"]).float()
    dense = torch.from_numpy(z[dense_key]).clamp(0, 1).transpose(0, 2).contiguous()[None, None]
    pooled = F.max_pool3d(dense, kernel_size=3, padding=1, stride=1)
    unsure = (pooled - thres).abs() < 1e-5
    assert vol.shape == ref.shape
    assert torch.equal(vol[~unsure], ref[~unsure]), f"{int((vol != ref)[~unsure].sum())} occupancy voxels differ"
    assert int((vol != ref).sum()) <= int(unsure.sum())
    return int((vol != ref).sum())


def _check_state(model, z, tag, exact=True):
    assert model.gridSize.tolist() == z[f"{tag}/gridSize"].tolist()
    assert torch.allclose(model.aabb.cpu(), torch.from_numpy(z[f"{tag}/aabb"]), rtol=0, atol=1e-6)
    assert abs(float(model.stepSize) - float(z[f"{tag}/stepSize"])) <= 1e-7
    assert model.nSamples == int(z[f"{tag}/nSamples"])
    sd = model.state_dict()
    for k in z.files:
        if k.startswith(f"{tag}/param/"):
            name = k[len(tag) + 7:]
            ref = torch.from_numpy(z[k])
            got = sd[name].detach().cpu()
            assert got.shape == ref.shape, (name, got.shape, ref.shape)
            if exact:
                assert torch.equal(got, ref), name
            else:
                assert float((got - ref).abs().max()) <= 2e-6 * max(1.0, float(ref.abs().max())), name


def test_maintenance_sequence_vs_reference(f3, cuda_device):
    z, spec, params = f3
    model = build_model(spec, params, cuda_device)
    with quiet():
        alpha, dense_xyz = model.getDenseAlpha((20, 24, 28))
        assert float((alpha.cpu() - torch.from_numpy(z["dense_alpha"])).abs().max()) <= 2e-6
        assert float((dense_xyz.cpu() - torch.from_numpy(z["dense_xyz"])).abs().max()) <= 1e-6
        new_aabb = model.updateAlphaMask((20, 24, 28))
    flips = _check_mask(model, z, "mask1", "dense_alpha")
    if flips == 0:
        assert torch.allclose(new_aabb.cpu(), torch.from_numpy(z["mask1/new_aabb"]), rtol=0, atol=1e-6)
    new_aabb = torch.from_numpy(z["mask1/new_aabb"]).to(cuda_device)        # continue from the reference's box
    if flips:
        model.alphaMask.alpha_volume.copy_(torch.from_numpy(z["mask1/volume"]).float().to(cuda_device))

    # ---- filtering_rays, both modes (tensorBase.py:372-404)
    rays = torch.from_numpy(z["filter/rays"])
    n = rays.shape[0]
    tag = torch.arange(n, dtype=torch.float32)[:, None].expand(n, 3).contiguous()
    depth = torch.arange(n, dtype=torch.float32)
    with quiet():
        r, c, d = model.filtering_rays(rays, tag, all_depth=depth, bbox_only=True)
    assert c[:, 0].long().tolist() == z["filter/bbox_keep"].tolist()
    assert torch.equal(r, rays[c[:, 0].long()]) and torch.equal(d, depth[c[:, 0].long()])
    with quiet():
        r, c = model.filtering_rays(rays, tag, N_samples=int(z["filter/alpha_n_samples"]), chunk=256, bbox_only=False)
    assert c[:, 0].long().tolist() == z["filter/alpha_keep"].tolist()
    assert torch.equal(r, rays[c[:, 0].long()])

    # ---- shrink -> upsample -> second mask update through the first mask -> shrink with the aabb correction
    with quiet():
        model.shrink(new_aabb)
    _check_state(model, z, "shrink1", exact=True)
    for p in list(model.density_plane) + list(model.app_plane):
        assert p.is_contiguous(memory_format=torch.channels_last)
    with quiet():
        model.upsample_volume_grid([30, 33, 37])
    _check_state(model, z, "up1", exact=False)
    for p in list(model.density_plane) + list(model.app_plane):
        assert p.is_contiguous(memory_format=torch.channels_last)
    # continue from the reference's upsampled factors so that later stages compare like with like
    model.load_state_dict({k[len("up1/param/"):]: torch.from_numpy(z[k]).to(cuda_device) for k in z.files
                           if k.startswith("up1/param/")})
    with quiet():
        alpha2, _ = model.getDenseAlpha((25, 26, 27))
        assert float((alpha2.cpu() - torch.from_numpy(z["dense_alpha2"])).abs().max()) <= 2e-6
        new_aabb2 = model.updateAlphaMask((25, 26, 27))
    flips2 = _check_mask(model, z, "mask2", "dense_alpha2")
    if flips2 == 0:
        assert torch.allclose(new_aabb2.cpu(), torch.from_numpy(z["mask2/new_aabb"]), rtol=0, atol=1e-6)
    with quiet():
        model.shrink(torch.from_numpy(z["mask2/new_aabb"]).to(cuda_device))
    _check_state(model, z, "shrink2", exact=True)

    # the maintained model still renders (the kernels take the new grid / box / mask)
    g = torch.Generator().manual_seed(0)
    d = torch.cat([0.3 * (torch.rand(256, 2, generator=g) * 2 - 1), torch.ones(256, 1)], -1)
    rays = torch.cat([torch.zeros(256, 3), d / d.norm(dim=-1, keepdim=True)], -1).to(cuda_device)
    with torch.no_grad():
        rgb, depth_map, zv, w = model(rays, is_train=False, white_bg=True, N_samples=-1)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    spec2 = orc.FieldSpec(**{**spec.__dict__, "aabb": model.aabb.cpu().tolist(), "grid": model.gridSize.tolist()})
    mask = (model.alphaMask.alpha_volume.detach().cpu(), model.alphaMask.aabb.cpu())
    ref = orc.render(spec2, sd, rays.cpu(), model.nSamples, False, True, None, mask)
    assert torch.equal(zv.cpu(), ref[2])
    assert float((w.cpu() - ref[3]).abs().max()) <= 2e-6
    assert float((rgb.cpu() - ref[0]).abs().max()) <= 1e-4


def test_checkpoint_roundtrip_with_packbits_mask(f3, cuda_device, tmp_path):
    """save() writes {'kwargs','state_dict','alphaMask.shape','alphaMask.mask' (np.packbits),'alphaMask.aabb'}
    (tensorBase.py:275-283); a second model built from ckpt['kwargs'] + load() renders identically.  The mask bits
    equal what the reference's own packbits of its volume gives."""
    from text2nerf_b200 import TensorVMSplit
    z, spec, params = f3
    model = build_model(spec, params, cuda_device)
    with quiet():
        model.updateAlphaMask((20, 24, 28))
    path = os.path.join(tmp_path, "ckpt.th")
    model.save(path)
    ckpt = torch.load(path, map_location=cuda_device, weights_only=False)
    assert set(ckpt.keys()) == {"kwargs", "state_dict", "alphaMask.shape", "alphaMask.mask", "alphaMask.aabb"}
    assert tuple(ckpt["alphaMask.shape"]) == tuple(z["mask1/volume"].shape)
    ours = model.alphaMask.alpha_volume.bool().cpu().numpy()
    assert np.array_equal(ckpt["alphaMask.mask"], np.packbits(ours.reshape(-1)))
    kwargs = ckpt["kwargs"]
    kwargs.update({"device": cuda_device})
    with quiet():
        twin = TensorVMSplit(**kwargs)
        twin.load(ckpt)
    assert torch.equal(twin.alphaMask.alpha_volume.cpu(), model.alphaMask.alpha_volume.cpu())
    g = torch.Generator().manual_seed(1)
    d = torch.cat([0.3 * (torch.rand(128, 2, generator=g) * 2 - 1), torch.ones(128, 1)], -1)
    rays = torch.cat([torch.zeros(128, 3), d / d.norm(dim=-1, keepdim=True)], -1).to(cuda_device)
    with torch.no_grad():
        a = model(rays, is_train=False, white_bg=True)
        b = twin(rays, is_train=False, white_bg=True)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
