"""Fused data loss: TensorBase.data_loss / t2n_data_loss / t2n_render_backward_tg against the
golden loss and gradients produced by the UNMODIFIED reference (text2nerf_main.py:556-575 composed from tensor ops),
and the loss kernel alone against a float64 restatement of the formula, NaN depth included."""
import ctypes as C

import pytest
import torch

from helpers import Case, build_model, cosine, fused_loss_with_jitter, golden_names, render_with_jitter, scaled_err
from oracle import t2n_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_names("train"))
def test_fused_loss_and_gradients_vs_golden(name, cuda_device):
    c = Case(name)
    model = build_model(c.spec, c.params, cuda_device, c.alpha)
    loss, l_rgb, l_depth, l_trans = fused_loss_with_jitter(model, c.rays.to(cuda_device), c.jitter, c.white_eff,
                                                           c.n_samples, c.rgb_gt, c.depth_gt)
    assert abs(float(loss) - c.loss) <= 2e-5 * abs(c.loss)
    assert abs(float(l_rgb + 0.005 * l_depth + 1e3 * l_trans) - float(loss)) <= 1e-6 * abs(float(loss))
    loss.backward()
    torch.cuda.synchronize()
    for k, p in model.named_parameters():
        g = p.grad
        assert g is not None, k
        assert scaled_err(g, c.grads[k]) <= 2e-4, (k, scaled_err(g, c.grads[k]))
        assert cosine(g, c.grads[k]) > 1 - 1e-6, k


def test_fused_equals_composed_path_and_scales_with_the_incoming_gradient(cuda_device):
    c = Case("t2n_noview_train")
    rays = c.rays.to(cuda_device)
    ref = build_model(c.spec, c.params, cuda_device, c.alpha)
    out = render_with_jitter(ref, rays, c.jitter, True, c.white_eff, c.n_samples)
    (3.0 * orc.training_loss(*out, c.rgb_gt.to(cuda_device), c.depth_gt.to(cuda_device))).backward()
    fused = build_model(c.spec, c.params, cuda_device, c.alpha)
    loss = fused_loss_with_jitter(fused, rays, c.jitter, c.white_eff, c.n_samples, c.rgb_gt, c.depth_gt)[0]
    (3.0 * loss).backward()
    torch.cuda.synchronize()
    for (k, a), (_, b) in zip(fused.named_parameters(), ref.named_parameters()):
        assert scaled_err(a.grad, b.grad) <= 2e-5, (k, scaled_err(a.grad, b.grad))
    # no_grad: value only, nothing saved
    with torch.no_grad():
        v = fused_loss_with_jitter(fused, rays, c.jitter, c.white_eff, c.n_samples, c.rgb_gt, c.depth_gt)[0]
    assert not v.requires_grad and abs(float(v) - float(loss)) <= 1e-6 * abs(float(loss))


def test_public_data_loss_draws_the_reference_rng_sequence(cuda_device):
    """data_loss() must consume the CPU generator exactly like forward(is_train=True): one U[0,1) per ray, plus one
    draw for the background only when white_bg is False (tensorBase.py:313-317, :497)."""
    c = Case("t2n_noview_train")
    model = build_model(c.spec, c.params, cuda_device, c.alpha)
    rays = c.rays.to(cuda_device)
    torch.manual_seed(11)
    a = model.data_loss(rays, c.rgb_gt, c.depth_gt, white_bg=True, N_samples=c.n_samples)
    after_a = torch.rand(1)
    torch.manual_seed(11)
    out = model(rays, is_train=True, white_bg=True, N_samples=c.n_samples)
    after_b = torch.rand(1)
    b = orc.training_loss(*out, c.rgb_gt.to(cuda_device), c.depth_gt.to(cuda_device))
    assert torch.equal(after_a, after_b)
    assert abs(float(a) - float(b)) <= 2e-6 * abs(float(b))


def test_loss_kernel_vs_float64_formula_with_nan_depth(cuda_device):
    from text2nerf_b200 import _native as nat
    lib = nat.load()
    g = torch.Generator().manual_seed(5)
    R, S = 257, 77
    rgb = torch.rand(R, 3, generator=g)
    depth = 1 + 5 * torch.rand(R, generator=g)
    depth[::9] = float("nan")
    z = torch.sort(0.5 + 7 * torch.rand(R, S, generator=g), dim=1).values
    w = torch.rand(R, S, generator=g) * 0.05
    rgb_gt, depth_gt = torch.rand(R, 3, generator=g), 1 + 5 * torch.rand(R, generator=g)
    w_depth, w_trans, delta, n_total = 0.005, 1e3, 0.1, 2 * R        # n_total: as one of two ray shards
    # float64 restatement with autograd (mask decided in float32 exactly like the tensor expression)
    rgb64, dep64, w64 = rgb.double().requires_grad_(), depth.double().requires_grad_(), w.double().requires_grad_()
    dep_fixed = torch.where(torch.isnan(dep64), torch.zeros_like(dep64), dep64)
    mask = ((z - depth_gt[:, None] + delta) < 0)
    mean_w = (w64 * mask).mean(1)
    terms = torch.stack([((rgb64 - rgb_gt.double()) ** 2).sum(1), (dep_fixed - depth_gt.double()) ** 2, mean_w ** 2], 1)
    loss = (terms[:, 0].sum() / 3 + w_depth * terms[:, 1].sum() + w_trans * terms[:, 2].sum()) / n_total
    loss.backward()
    dev = cuda_device
    t = {k: v.to(dev).contiguous() for k, v in dict(rgb=rgb, depth=depth, z=z, w=w, rgb_gt=rgb_gt, depth_gt=depth_gt).items()}
    o = {k: torch.empty(s, device=dev) for k, s in dict(terms=(R, 3), g_rgb=(R, 3), g_depth=(R,), coef=(R,), g_w=(R, S)).items()}
    outs = nat.T2NOutputs(t["rgb"].data_ptr(), t["depth"].data_ptr(), t["z"].data_ptr(), t["w"].data_ptr())
    rc = lib.t2n_data_loss(C.byref(outs), R, S, t["rgb_gt"].data_ptr(), t["depth_gt"].data_ptr(), w_depth, w_trans, delta,
                           1.0 / n_total, o["terms"].data_ptr(), o["g_rgb"].data_ptr(), o["g_depth"].data_ptr(),
                           o["coef"].data_ptr(), o["g_w"].data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
    nat.check(rc, "t2n_data_loss")
    torch.cuda.synchronize()
    assert scaled_err(o["terms"], terms.detach()) <= 2e-6
    assert scaled_err(o["g_rgb"], rgb64.grad) <= 2e-6
    gd = torch.nan_to_num(dep64.grad, nan=0.0)
    assert scaled_err(o["g_depth"], gd) <= 2e-6
    assert float(o["g_depth"].cpu()[::9].abs().max()) == 0.0        # NaN depth -> replaced by a constant -> zero gradient
    assert scaled_err(o["g_w"], w64.grad) <= 2e-6
    # compact form == dense form
    dense = o["coef"][:, None] * ((t["z"] - t["depth_gt"][:, None] + delta) < 0)
    assert torch.equal(dense, o["g_w"])
    # bad arguments are refused
    assert lib.t2n_data_loss(None, R, S, None, None, 0.0, 0.0, 0.0, 1.0, None, None, None, None, None, None) != 0
