"""Pin the oracle: it must reproduce every golden vector generated from the unmodified
reference (tests/golden/make_golden.py).  Forward outputs bit-for-bit; gradients of the
Text2NeRF data loss through torch autograd of the oracle to float round-off."""
import numpy as np
import pytest
import torch

from helpers import Case, golden_names
from oracle import t2n_oracle as orc


@pytest.mark.parametrize("name", golden_names())
def test_forward_bit_exact(name):
    c = Case(name)
    torch.set_num_threads(1)
    rgb, depth, z, w = orc.render(c.spec, c.params, c.rays, c.n_samples, c.is_train, c.white_eff, c.jitter, c.alpha)
    assert torch.equal(z, c.out["z_vals"])
    assert torch.equal(w, c.out["weight"])
    assert torch.equal(rgb, c.out["rgb_map"])
    assert torch.equal(depth, c.out["depth_map"])


@pytest.mark.parametrize("name", golden_names("train"))
def test_gradients(name):
    c = Case(name)
    torch.set_num_threads(1)
    params = {k: v.clone().requires_grad_(True) for k, v in c.params.items()}
    out = orc.render(c.spec, params, c.rays, c.n_samples, True, c.white_eff, c.jitter, c.alpha)
    loss = orc.training_loss(*out, c.rgb_gt, c.depth_gt)
    assert abs(float(loss) - c.loss) <= 1e-6 * abs(c.loss)
    loss.backward()
    for k, g_ref in c.grads.items():
        g = params[k].grad
        g = torch.zeros_like(g_ref) if g is None else g
        scale = float(g_ref.abs().max())
        assert float((g - g_ref).abs().max()) <= 1e-5 * max(scale, 1e-12), k


def test_indexform_bilinear_matches_grid_sample():
    """The explicit-tap restatement (the arithmetic the CUDA kernels implement) against ATen."""
    c = Case("t2n_noview_eval")
    _, _, _, _, aux = orc.render(c.spec, c.params, c.rays, c.n_samples, False, True, None, None, keep=True)
    xn = aux["xn"][aux["valid"]]
    a = orc.density_feature(c.params, xn)
    b = orc.density_feature_indexform(c.params, xn, c.spec.grid)
    assert float((a - b).abs().max()) < 2e-5 * float(a.abs().max())


def test_step_and_samples_match_fixture_shapes():
    for name in golden_names():
        c = Case(name)
        S = c.n_samples if c.n_samples > 0 else orc.derive_step(c.spec)[1]
        assert c.out["z_vals"].shape == (c.rays.shape[0], S)


def test_get_rays_golden():
    z = np.load(__import__("os").path.join(__import__("helpers").GOLDEN_DIR, "get_rays.npz"))
    H, W = int(z["H"]), int(z["W"])
    dirs = orc.pixel_directions(H, W, list(z["focal"]))
    assert torch.equal(dirs, torch.from_numpy(z["directions"]))
    dn = dirs / torch.norm(dirs, dim=-1, keepdim=True)
    ro, rd = orc.camera_rays(dn, torch.from_numpy(z["c2w"]))
    assert torch.equal(ro, torch.from_numpy(z["rays_o"])) and torch.equal(rd, torch.from_numpy(z["rays_d"]))
