"""CPU-side checks: the C-ABI library loads and exports every symbol include/t2n_b200.h
declares (no compute calls without a GPU), argument validation, and the host logic of the
Python mirror (parameter layout, decoder column recipe, step size, state-dict surface)."""
import ctypes as C
import os
import re

import pytest
import torch

from helpers import Case, build_model, quiet
from oracle import t2n_oracle as orc
from text2nerf_b200 import _native as nat
from text2nerf_b200.tensorBase import decoder_recipe

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = nat.load()
    header = open(os.path.join(ROOT, "include", "t2n_b200.h")).read()
    declared = set(re.findall(r"\b(t2n_[a-z_0-9]+)\s*\(", header))
    assert declared == set(nat.SYMBOLS), declared ^ set(nat.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.t2n_abi_version() == nat.ABI_VERSION
    assert b"success" in lib.t2n_error_string(0)


def test_struct_sizes_match_header_layout():
    # natural alignment, no packing surprises: sizes are what the C compiler produces for the header
    assert C.sizeof(nat.T2NField) == 4 * (9 + 3 + 1 + 2 + 4 + 4 + 6 + 2 + 2)
    assert C.sizeof(nat.T2NParams) == 8 * (12 + 1 + 6 + 2)
    assert C.sizeof(nat.T2NGrads) == 8 * (12 + 1 + 6 + 1)       # + app_done_event
    assert C.sizeof(nat.T2NBatch) == 8 * 2 + 4 * 4
    assert C.sizeof(nat.T2NScratch) == 8 * 19
    assert C.sizeof(nat.T2NAlphaMask) == 8 + 4 * 3 + 4 * 6 + 4   # padded to 8


def test_argument_validation_without_gpu():
    lib = nat.load()
    f, p = nat.T2NField(), nat.T2NParams()
    assert lib.t2n_render_forward(None, None, None, None, None, None, None) == -1
    f.grid = nat.I3(8, 8, 8)
    f.n_sigma = nat.I3(6, 6, 6)          # not a multiple of 4
    f.n_app = nat.I3(8, 8, 8)
    assert lib.t2n_render_forward(C.byref(f), C.byref(p), None, None, None, None, None) == -2
    assert b"layout" in lib.t2n_error_string(-2)
    assert lib.t2n_get_rays(None, 1.0, 1.0, 0.0, 0.0, 4, 4, 0, None, None) == -1


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(nat.NativeLibraryError):
        nat.load(str(tmp_path / "nope.so"))


def test_forward_refuses_cpu_tensors():
    c = Case("t2n_noview_eval")
    m = build_model(c.spec, c.params, "cpu")
    with pytest.raises(nat.NativeLibraryError):
        m(c.rays, is_train=False, white_bg=True, N_samples=8)


@pytest.mark.parametrize("mode,fea_pe,view_pe", [("MLP_Fea_noview", 6, 2), ("MLP_Fea", 2, 2), ("MLP", 6, 6),
                                                 ("MLP_Fea_noview", 0, 0), ("MLP_Fea", 3, 0)])
def test_decoder_recipe_reproduces_reference_columns(mode, fea_pe, view_pe):
    """Evaluate the pair recipe in numpy-like torch code and compare against the oracle's concat
    of [features | viewdirs | PE...] after un-permuting: every reference column must be produced
    exactly once and with the same value."""
    A = 27
    spec = orc.FieldSpec(aabb=[[-1, -1, -1], [1, 1, 1]], grid=[8, 8, 8], shading=mode, fea_pe=fea_pe, view_pe=view_pe)
    mlp_in, perm, pairs = decoder_recipe(mode, A, fea_pe, view_pe)
    assert mlp_in == orc.mlp_in_dim(spec)
    assert len(perm) % 32 == 0 and len(pairs) * 2 == len(perm)
    used = [k for k in perm if k >= 0]
    assert sorted(used) == list(range(mlp_in))
    g = torch.Generator().manual_seed(0)
    feat, view, xn = torch.randn(5, A, generator=g), torch.randn(5, 3, generator=g), torch.randn(5, 3, generator=g)
    base = torch.cat([feat, view, xn, torch.zeros(5, 1)], -1)
    cols = []
    for d in pairs:
        sa, sb, f, trig = d & 0xff, (d >> 8) & 0xff, (d >> 16) & 0xf, (d >> 20) & 1
        if trig:
            v = base[:, sa] * float(1 << f)
            cols += [torch.sin(v), torch.cos(v)]
        else:
            cols += [base[:, sa], base[:, sb]]
    internal = torch.stack(cols, -1)
    # reference order
    ref_cols = [feat] + ([view] if mode != "MLP_Fea_noview" else [])
    if mode in ("MLP_Fea_noview", "MLP_Fea") and fea_pe > 0:
        ref_cols.append(orc.freq_encode(feat, fea_pe))
    if mode in ("MLP_Fea", "MLP") and view_pe > 0:
        ref_cols.append(orc.freq_encode(view, view_pe))
    ref = torch.cat(ref_cols, -1)
    for k, src in enumerate(perm):
        if src >= 0:
            assert torch.equal(internal[:, k], ref[:, src]), (k, src)
        else:
            assert float(internal[:, k].abs().max()) == 0.0


def test_module_surface_matches_reference_state_dict():
    c = Case("t2n_noview_train")
    m = build_model(c.spec, c.params, "cpu")
    sd = m.state_dict()
    # key order of the reference's state_dict
    order = [f"{g}.{i}" for g in ("density_plane", "density_line", "app_plane", "app_line") for i in range(3)]
    order += ["basis_mat.weight"] + [f"renderModule.mlp.{l}.{w}" for l in (0, 2, 4) for w in ("weight", "bias")]
    assert list(sd.keys()) == order
    assert set(sd.keys()) == set(c.params.keys())
    for k, v in c.params.items():
        assert sd[k].shape == v.shape and torch.equal(sd[k], v)
    # planes/lines are physically texel-major while keeping the reference's logical shape
    for p in list(m.density_plane) + list(m.app_plane) + list(m.density_line) + list(m.app_line):
        assert p.is_contiguous(memory_format=torch.channels_last)
    groups = m.get_optparam_groups(0.02, 1e-3)
    assert [g["lr"] for g in groups] == [0.02, 0.02, 0.02, 0.02, 1e-3, 1e-3]
    kw = m.get_kwargs()
    assert kw["gridSize"] == list(c.spec.grid) and kw["shadingMode"] == "MLP_Fea_noview"
    step, n = orc.derive_step(c.spec)
    assert float(m.stepSize) == float(step) and m.nSamples == n


def test_save_load_roundtrip(tmp_path):
    from text2nerf_b200 import TensorVMSplit
    c = Case("lego_relu_alphamask_train")
    m = build_model(c.spec, c.params, "cpu", c.alpha)
    path = str(tmp_path / "ck.th")
    m.save(path)
    ck = torch.load(path, weights_only=False)
    kw = ck["kwargs"]
    kw.update({"device": "cpu"})
    with quiet():
        m2 = TensorVMSplit(**kw)
    m2.load(ck)
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k
    assert torch.equal(m2.alphaMask.alpha_volume.bool(), m.alphaMask.alpha_volume.bool())


def test_flat_grad_buffer_layout():
    c = Case("t2n_noview_train")
    m = build_model(c.spec, c.params, "cpu")
    buf = m.enable_flat_grads(True)
    views = m._flat_grad["views"]
    params = m._flat_params()
    assert len(views) == len(params) == 19
    assert buf.numel() >= sum(p.numel() for p in params)
    for v, p in zip(views, params):
        assert v.shape == p.shape and v.stride() == p.stride()
        assert v.data_ptr() % 16 == 0
    views[0].fill_(1.0)
    assert float(buf.sum()) == params[0].numel()


def test_tv_regulariser_matches_oracle():
    c = Case("t2n_noview_train")
    m = build_model(c.spec, c.params, "cpu")
    ref = sum(orc.tv_plane(c.params[f"density_plane.{i}"]) for i in range(3)) * 1e-2
    got = m.TV_loss_density(orc.tv_plane)
    assert abs(float(got) - float(ref)) <= 1e-6 * abs(float(ref))


def test_training_entry_points_validate_arguments_without_gpu():
    """The fused-loss / TV / Adam entry points refuse bad arguments before touching a device."""
    lib = nat.load()
    assert lib.t2n_data_loss(None, 4, 4, None, None, 0.0, 0.0, 0.0, 1.0, None, None, None, None, None, None) == -1
    assert lib.t2n_tv_plane_sums(None, 4, 4, 4, None, None) == -1
    assert lib.t2n_tv_plane_grad(None, 4, 4, 4, None, 1.0, 1.0, None, None) == -1
    assert lib.t2n_adam_step(None, 0, 0.9, 0.99, 1e-8, 0.0, 1, None) == -1
    t = (nat.T2NAdamTensor * 1)()
    assert lib.t2n_adam_step(t, 1, 1.5, 0.99, 1e-8, 0.0, 1, None) == -1       # beta1 out of range
    assert lib.t2n_adam_step(t, 1, 0.9, 0.99, 1e-8, 0.0, 0, None) == -1       # step counts from 1
    assert C.sizeof(nat.T2NTransGrad) == 8 * 3 and C.sizeof(nat.T2NAdamTensor) == 8 * 6


def test_fused_training_paths_refuse_cpu_tensors_and_generic_tv_callables_stay_tensor_ops():
    from text2nerf_b200.optim import FusedAdam
    c = Case("t2n_noview_train")
    m = build_model(c.spec, c.params, "cpu")
    with pytest.raises(nat.NativeLibraryError):
        m.data_loss(c.rays, c.rgb_gt, c.depth_gt, N_samples=c.n_samples)
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    with pytest.raises(nat.NativeLibraryError):
        FusedAdam([p]).step()
    with pytest.raises(ValueError):
        FusedAdam([p], betas=(1.0, 0.99))
    # TV_loss_* take arbitrary callables (tensoRF.py:193-203): CPU planes / non-TVLoss callables use tensor ops
    class TVLoss(torch.nn.Module):
        TVLoss_weight = 1

        def forward(self, x):
            return orc.tv_plane(x)
    want = sum(orc.tv_plane(pl) * 1e-2 for pl in m.density_plane)
    assert torch.equal(m.TV_loss_density(TVLoss()), want)                 # CPU planes: no kernel route
    assert torch.equal(m.TV_loss_density(orc.tv_plane), want)
    assert torch.equal(m.TV_loss_app(lambda x: x.abs().mean()), sum(pl.abs().mean() * 1e-2 for pl in m.app_plane))


def test_tf32_round_bit_pattern_is_round_to_nearest_ties_away():
    """csrc/operand_image.cuh tf32_hi: (bits + 0x1000) & 0xffffe000 is cvt.rna.tf32.f32 (round to nearest, ties away from
    zero, onto 10 explicit mantissa bits) for every finite input; checked against exact rational arithmetic."""
    import numpy as np
    from fractions import Fraction
    rng = np.random.default_rng(0)
    bits = rng.integers(0, 2 ** 32, size=20000, dtype=np.uint64).astype(np.uint32)
    # add the interesting patterns: ties, all-ones mantissas (carry into the exponent), tiny and huge exponents
    extra = [0x3f800000 | 0x1000, 0x3f800000 | 0x0fff, 0x3f800000 | 0x1001, 0x3fffffff, 0x3f7ff000, 0xbf801000, 0x00801000,
             0x7e7ff000, 0x3f803000, 0xbf802fff]
    bits = np.concatenate([bits, np.array(extra, dtype=np.uint32)])
    expo = (bits >> 23) & 0xff
    bits = bits[(expo != 0) & (expo != 255)]                       # normal numbers (the kernels never split inf / nan / denormals)
    got = ((bits.astype(np.uint64) + 0x1000) & 0xffffe000).astype(np.uint32)
    for b, g in zip(bits.tolist()[:3000] + bits.tolist()[-len(extra):], got.tolist()[:3000] + got.tolist()[-len(extra):]):
        x = Fraction(float(np.array([b], dtype=np.uint32).view(np.float32)[0]))
        e = ((b >> 23) & 0xff) - 127
        ulp = Fraction(2) ** (e - 10)                               # TF32 keeps 10 mantissa bits
        qf = abs(x) / ulp
        n = qf.numerator // qf.denominator
        if qf - n >= Fraction(1, 2):
            n += 1                                                  # ties away from zero
        want = n * ulp * (-1 if x < 0 else 1)
        gv = np.array([g], dtype=np.uint32).view(np.float32)[0]
        if np.isfinite(gv):
            assert Fraction(float(gv)) == want, hex(b)
        assert g & 0x1fff == 0
