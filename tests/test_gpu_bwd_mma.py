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
    assert err < 2e-6, err
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
    """A batch that lists more samples than the operand images can hold takes the recomputing FFMA kernel;
    both paths must give the same gradients (and the golden ones, checked in test_gpu_parity)."""
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
        res[tag] = (_grads(model), listed)
    assert res["mma"][1] == res["ffma"][1] > 128, "the case must overflow a 128-row capacity"
    for k, g in res["mma"][0].items():
        g2 = res["ffma"][0][k]
        if float(g2.abs().max()) == 0.0:
            assert float(g.abs().max()) == 0.0, k
            continue
        assert scaled_err(g, g2) <= 2e-4, (k, scaled_err(g, g2))
        assert cosine(g, g2) > 1 - 1e-6, k
