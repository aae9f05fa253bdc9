"""Fused TV regulariser: TensorVMSplit.TV_loss_density / TV_loss_app with the reference's
utils.TVLoss as `reg` run t2n_tv_plane_sums / t2n_tv_plane_grad; value and gradients must equal the tensor-op
formulation (oracle.tv_plane = utils.py:488-504 restated, checked against the live reference in test_host_cpu /
test_oracle_live)."""
import pytest
import torch

from helpers import Case, build_model, scaled_err
from oracle import t2n_oracle as orc

pytestmark = pytest.mark.gpu


class TVLoss(torch.nn.Module):
    """Stand-in with the reference class's name and attribute (utils.py:488-504): what text2nerf_main.py passes as `tvreg`."""

    def __init__(self, TVLoss_weight=1):
        super().__init__()
        self.TVLoss_weight = TVLoss_weight

    def forward(self, x):
        return self.TVLoss_weight * orc.tv_plane(x)


def _models(cuda_device, name="lego_relu_alphamask_train"):
    c = Case(name)
    return build_model(c.spec, c.params, cuda_device, c.alpha), build_model(c.spec, c.params, cuda_device, c.alpha)


@pytest.mark.parametrize("weight", [1.0, 0.37])
def test_fused_tv_matches_tensor_ops(weight, cuda_device):
    fused, ref = _models(cuda_device)
    reg = TVLoss(weight)
    lf = fused.TV_loss_density(reg) * 0.1 + fused.TV_loss_app(reg) * 0.01
    lr = ref.TV_loss_density(lambda x: weight * orc.tv_plane(x)) * 0.1 + ref.TV_loss_app(lambda x: weight * orc.tv_plane(x)) * 0.01
    assert abs(float(lf) - float(lr)) <= 1e-5 * abs(float(lr))
    lf.backward()
    lr.backward()
    torch.cuda.synchronize()
    for (k, a), (_, b) in zip(fused.named_parameters(), ref.named_parameters()):
        if "plane" in k:
            assert a.grad is not None and scaled_err(a.grad, b.grad) <= 1e-5, (k, scaled_err(a.grad, b.grad))
        else:
            assert a.grad is None and b.grad is None, k


def test_fused_tv_accumulates_into_the_flat_gradient_buffer(cuda_device):
    fused, ref = _models(cuda_device, "t2n_noview_train")
    flat = fused.enable_flat_grads(True)
    flat.zero_()
    reg = TVLoss()
    (fused.TV_loss_density(reg) * 0.1).backward()
    (fused.TV_loss_density(reg) * 0.1).backward()           # second call accumulates
    (ref.TV_loss_density(reg.forward) * 0.2).backward()
    torch.cuda.synchronize()
    views = fused._flat_grad["views"]
    for i, p in enumerate(ref.density_plane):
        assert scaled_err(views[i], p.grad) <= 1e-5
    assert all(p.grad is None for p in fused.density_plane)  # flat mode: autograd gets nothing
    assert float(views[6].abs().max()) == 0.0                # app planes untouched


def test_tv_kernels_on_a_non_square_plane(cuda_device):
    import ctypes as C
    from text2nerf_b200 import _native as nat
    lib = nat.load()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1, 12, 37, 53, generator=g).to(cuda_device).contiguous(memory_format=torch.channels_last)
    blocks = int(lib.t2n_tv_blocks())
    partials = torch.zeros((blocks, 2), device=cuda_device)
    st = torch.cuda.current_stream(cuda_device).cuda_stream
    nat.check(lib.t2n_tv_plane_sums(x.data_ptr(), 37, 53, 12, partials.data_ptr(), st), "sums")
    xr = x.detach().clone().double().requires_grad_()
    h_tv = ((xr[:, :, 1:, :] - xr[:, :, :-1, :]) ** 2).sum()
    w_tv = ((xr[:, :, :, 1:] - xr[:, :, :, :-1]) ** 2).sum()
    s = partials.sum(0).double().cpu()
    assert abs(float(s[0]) - float(h_tv)) <= 2e-6 * float(h_tv) and abs(float(s[1]) - float(w_tv)) <= 2e-6 * float(w_tv)
    (0.7 * (0.3 * h_tv + 1.9 * w_tv)).backward()
    grad = torch.zeros_like(x)
    gout = torch.tensor([0.7], device=cuda_device)
    nat.check(lib.t2n_tv_plane_grad(x.data_ptr(), 37, 53, 12, gout.data_ptr(), 0.3, 1.9, grad.data_ptr(), st), "grad")
    torch.cuda.synchronize()
    assert scaled_err(grad, xr.grad) <= 2e-6
    assert lib.t2n_tv_plane_sums(x.data_ptr(), 37, 53, 10, partials.data_ptr(), st) != 0      # C % 4 != 0 is refused
