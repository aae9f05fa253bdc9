"""CPU tests of the small surface functions around the hot path against goldens the UNMODIFIED reference produced
(tests/golden/make_golden_f3.py): models/sh.py eval_sh / eval_sh_bases for degrees 0-4, density_L1 and
vector_comp_diffs (models/tensoRF.py:173-191), and the host-side fixes of the round-1 advisor findings."""
import copy
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN_DIR, build_model, quiet
from oracle import t2n_oracle as orc
from text2nerf_b200 import sh


@pytest.fixture(scope="module")
def shz():
    return np.load(os.path.join(GOLDEN_DIR, "sh_eval.npz"))


@pytest.mark.parametrize("deg", [0, 1, 2, 3, 4])
def test_eval_sh_bases_vs_reference(deg, shz):
    dirs = torch.from_numpy(shz["dirs"])
    got = sh.eval_sh_bases(deg, dirs)
    ref = torch.from_numpy(shz[f"bases/{deg}"])
    assert got.shape == ref.shape
    assert float((got - ref).abs().max()) <= 1e-6


@pytest.mark.parametrize("deg", [0, 1, 2, 3, 4])
def test_eval_sh_vs_reference(deg, shz):
    dirs = torch.from_numpy(shz["dirs"])
    coeff = torch.from_numpy(shz[f"coeff/{deg}"])
    got = sh.eval_sh(deg, coeff, dirs)
    ref = torch.from_numpy(shz[f"eval/{deg}"])
    assert got.shape == ref.shape
    assert float((got - ref).abs().max()) <= 2e-6 * max(1.0, float(ref.abs().max()))


def _f3_model(device="cpu"):
    z = np.load(os.path.join(GOLDEN_DIR, "f3_maintenance.npz"))
    import json
    d = json.loads(str(z["spec"]))
    d.pop("dtype")
    spec = orc.FieldSpec(**d)
    params = {k[6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param/")}
    return z, spec, build_model(spec, params, device)


def test_density_l1_and_vector_diffs_vs_reference():
    z, _, model = _f3_model()
    assert abs(float(model.density_L1()) - float(z["reg/density_L1"])) <= 1e-6 * abs(float(z["reg/density_L1"]))
    assert abs(float(model.vector_comp_diffs()) - float(z["reg/vector_comp_diffs"])) <= 1e-5 * abs(float(z["reg/vector_comp_diffs"]))


def test_flat_grads_follow_upsampling():
    """the flat gradient views must be rebuilt when upsample_volume_grid / shrink replace the factor
    Parameters; stale views are refused instead of being scattered into out of bounds."""
    _, _, model = _f3_model()
    buf = model.enable_flat_grads(True)
    n0 = buf.numel()
    with quiet():
        model.upsample_volume_grid([24, 28, 32])
    flat = model._flat_grad
    assert flat["buffer"].numel() > n0
    p_cl = model._native_param_tensors(model._flat_params())
    views = model._grad_buffers(p_cl)
    for v, t in zip(views, p_cl):
        assert v.shape == t.shape and v.stride() == t.stride()
    # a replaced parameter without a refresh is caught
    model.density_plane[0] = torch.nn.Parameter(torch.zeros(1, 16, 9, 9).contiguous(memory_format=torch.channels_last))
    with pytest.raises(RuntimeError, match="stale"):
        model._grad_buffers(model._native_param_tensors(model._flat_params()))


def test_rays_wider_than_six_columns_are_refused():
    from text2nerf_b200.tensorBase import TensorBase
    with pytest.raises(ValueError, match="6"):
        TensorBase._check_rays(torch.zeros(4, 8))
    with pytest.raises(ValueError):
        TensorBase._check_rays(torch.zeros(4, 5))
    r = TensorBase._check_rays(torch.zeros(4, 6, dtype=torch.float64))
    assert r.dtype == torch.float32 and r.is_contiguous()


def test_module_deepcopy_with_poll_state():
    from text2nerf_b200.tensorBase import _ListedCountPoll
    _, _, model = _f3_model()
    model.__dict__["_listed_side"] = _ListedCountPoll()
    with quiet():
        twin = copy.deepcopy(model)
    assert twin.__dict__["_listed_side"] is not model.__dict__["_listed_side"]
    assert twin.__dict__["_listed_side"].pending == []
