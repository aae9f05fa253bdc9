"""GPU tests of the tensor-core backward: the generic weight-gradient GEMM kernel against a float64 matmul,
and the capacity-overflow fallback (FFMA kernel) against the tensor-core path on the same batch."""
import ctypes as C

import pytest
import torch

from helpers import Case, build_model, cosine, render_with_jitter, scaled_err
from oracle import t2n_oracle as orc

pytestmark = pytest.mark.gpu


def _image(lib, rows, ng, dev):
    n = rows.shape[0]
    tiles = (n + 127) // 128
    img = torch.zeros((tiles * 32768 * ng,), dtype=torch.uint8, device=dev)
    rc = lib.t2n_debug_make_image(rows.data_ptr(), n, ng, img.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
    assert rc == 0
    return img


@pytest.mark.parametrize("ngx,ngy,n", [(4, 4, 128), (4, 13, 1000), (1, 5, 3333), (4, 1, 77), (4, 8, 128 * 400)])
def test_wgrad_gemm_vs_float64(ngx, ngy, n, cuda_device):
    """out[i][k] = sum_m X[m][i] Y[m][k] with 3xTF32 on tcgen05 (MN-major operands) must be fp32-accurate."""
    from text2nerf_b200 import _native as nat
    lib = nat.load()
    g = torch.Generator().manual_seed(ngx * 100 + ngy)
    X = torch.randn(n, 32 * ngx, generator=g)
    Y = torch.randn(n, 32 * ngy, generator=g)
    Xd, Yd = X.to(cuda_device), Y.to(cuda_device)
    xi, yi = _image(lib, Xd, ngx, cuda_device), _image(lib, Yd, ngy, cuda_device)
    out = torch.zeros((128, 32 * ngy), device=cuda_device)
    ones = torch.zeros((128,), device=cuda_device)
    cnt = torch.tensor([n, 0, 0, 0], dtype=torch.int32, device=cuda_device)
    rc = lib.t2n_debug_wgrad(xi.data_ptr(), ngx, yi.data_ptr(), ngy, cnt.data_ptr(), 1 << 30, out.data_ptr(),
                             ones.data_ptr(), torch.cuda.current_stream(cuda_device).cuda_stream)
    nat.check(rc, "t2n_debug_wgrad")
    torch.cuda.synchronize()
    ref = X.double().t() @ Y.double()
    rows = 32 * ngx
    scale = float(ref.abs().max())
    err = float((out[:rows].cpu().double() - ref).abs().max()) / scale
    assert err < 2e-6 * max(1.0, (n / 4096) ** 0.5), err
    ref1 = X.double().sum(0)
    err1 = float((ones[:rows].cpu().double() - ref1).abs().max()) / float(ref1.abs().max())
    assert err1 < 2e-6, err1
    # capacity overflow: the kernel must leave the outputs untouched
    out.zero_()
    rc = lib.t2n_debug_wgrad(xi.data_ptr(), ngx, yi.data_ptr(), ngy, cnt.data_ptr(), n - 1, out.data_ptr(),
                             None, torch.cuda.current_stream(cuda_device).cuda_stream)
    assert rc == 0
    assert float(out.abs().max()) == 0.0


def _grads(model):
    return {k: p.grad.detach().clone() for k, p in model.named_parameters()}


@pytest.mark.parametrize("name", ["t2n_noview_train", "lego_mlp_fea_train", "lego_mlp_train"])
def test_overflow_fallback_matches_tensor_core_path(name, cuda_device, monkeypatch):
    """A batch that lists more samples than the operand images can hold is split: the tensor-core backward handles the
    rows that fit (here one 128-row tile), the recomputing FFMA kernel the rest; the sum must equal the all-tensor-core
    gradients (and the golden ones, checked in test_gpu_parity).  Heads outside the tensor-core envelope take the FFMA
    kernel for everything."""
    c = Case(name)
    res = {}
    for tag, cap in (("mma", None), ("ffma", "128")):
        if cap is None:
            monkeypatch.delenv("T2N_ACT_ROWS_MAX", raising=False)
        else:
            monkeypatch.setenv("T2N_ACT_ROWS_MAX", cap)
        model = build_model(c.spec, c.params, cuda_device, c.alpha)
        out = render_with_jitter(model, c.rays.to(cuda_device), c.jitter, True, c.white_eff, c.n_samples)
        listed = model.app_sample_count()[0]
        orc.training_loss(*out, c.rgb_gt.to(cuda_device), c.depth_gt.to(cuda_device)).backward()
        cnt = model._last_counters.cpu()
        from text2nerf_b200 import _native as nat
        in_envelope = nat.load().t2n_bwd_pack_floats(C.byref(model._native_field())) > 0
        if tag == "mma" and in_envelope:
            assert int(cnt[2]) == (listed + 127) // 128 and int(cnt[3]) == 0, cnt
        elif in_envelope:
            assert int(cnt[2]) == 1 and int(cnt[3]) > 0, cnt        # one tile on the tensor cores, the overflow on FFMA
        else:
            assert int(cnt[2]) == 0 and int(cnt[3]) > 0, cnt
        res[tag] = (_grads(model), listed)
    assert res["mma"][1] == res["ffma"][1] > 128, "the case must overflow a 128-row capacity"
    for k, g in res["mma"][0].items():
        g2 = res["ffma"][0][k]
        if float(g2.abs().max()) == 0.0:
            assert float(g.abs().max()) == 0.0, k
            continue
        assert scaled_err(g, g2) <= 2e-4, (k, scaled_err(g, g2))
        assert cosine(g, g2) > 1 - 1e-6, k


def test_tensor_core_backward_covers_view_dependent_heads(cuda_device):
    """MLP_Fea / MLP heads (view-direction columns, two PE groups) inside the tensor-core envelope (featureC 128,
    16-multiple component counts) against oracle autograd."""
    for shading, fea_pe, view_pe in (("MLP_Fea", 2, 2), ("MLP", 6, 4), ("MLP_Fea_noview", 6, 2)):
        spec = orc.FieldSpec(aabb=[[-8, -8, -8], [8, 8, 8]], grid=[40, 44, 48], near_far=[0.5, 8.0], step_ratio=1.0,
                             shading=shading, fea_pe=fea_pe, view_pe=view_pe)
        params = orc.init_params(spec, seed=11, density_gain=10.8, app_gain=3.0)
        g = torch.Generator().manual_seed(5)
        R = 300
        d = torch.cat([0.5 * (torch.rand(R, 2, generator=g) * 2 - 1), torch.ones(R, 1)], -1)
        rays = torch.cat([0.05 * torch.randn(R, 3, generator=g), d / d.norm(dim=-1, keepdim=True)], -1)
        jitter = torch.rand(R, 1, generator=g)
        S = orc.derive_step(spec)[1] // 2
        rgb_gt, depth_gt = torch.rand(R, 3, generator=g), 0.5 + 7 * torch.rand(R, generator=g)
        p_ref = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        orc.training_loss(*orc.render(spec, p_ref, rays, S, True, True, jitter), rgb_gt, depth_gt).backward()
        model = build_model(spec, params, cuda_device)
        out = render_with_jitter(model, rays.to(cuda_device), jitter, True, True, S)
        orc.training_loss(*out, rgb_gt.to(cuda_device), depth_gt.to(cuda_device)).backward()
        cnt = model._last_counters.cpu()
        assert int(cnt[2]) == (int(cnt[0]) + 127) // 128 > 0 and int(cnt[3]) == 0, (shading, cnt)
        for k, p in model.named_parameters():
            gr = p_ref[k].grad
            assert scaled_err(p.grad, gr) <= 2e-4, (shading, k, scaled_err(p.grad, gr))
            assert cosine(p.grad, gr) > 1 - 1e-6, (shading, k)
