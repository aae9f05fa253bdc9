"""Runs the UNMODIFIED reference modules placed under ``oracle/_ref`` by ``oracle/make_ref.py`` (CPU, PyTorch).

TEST / MEASUREMENT INFRASTRUCTURE, NOT PRODUCT CODE.  Used by bench.py's ``cpu_baseline`` leg and ``--impl reference``
(``kind: "reference"``) and by tests that pin the oracle port against it.
"""
import contextlib
import importlib
import io
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "MANIFEST.json"))


def modules():
    """(models.tensoRF, models.tensorBase) of the reference, imported from oracle/_ref."""
    if not available():
        raise RuntimeError("oracle/_ref is not built (python oracle/make_ref.py needs the reference checkout)")
    # a `models` package imported earlier from the reference checkout itself (tests in the authoring container) is the
    # same unmodified code: reuse it rather than importing the package twice
    if "models.tensoRF" not in sys.modules and REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    return importlib.import_module("models.tensoRF"), importlib.import_module("models.tensorBase")


def build(spec, params):
    """The reference's TensorVMSplit on the CPU with the given state (models/tensoRF.py:139, tensorBase.py:164-198)."""
    tensoRF, _ = modules()
    with contextlib.redirect_stdout(io.StringIO()):
        m = tensoRF.TensorVMSplit(
            spec.aabb_t(), list(spec.grid), "cpu",
            density_n_comp=list(spec.density_n_comp), appearance_n_comp=list(spec.app_n_comp),
            app_dim=spec.app_dim, near_far=list(spec.near_far), shadingMode=spec.shading,
            alphaMask_thres=0.001, density_shift=spec.density_shift, distance_scale=spec.distance_scale,
            pos_pe=spec.pos_pe, view_pe=spec.view_pe, fea_pe=spec.fea_pe, featureC=spec.featureC,
            step_ratio=spec.step_ratio, fea2denseAct=spec.act)
    m.load_state_dict({k: v.detach().clone() for k, v in params.items()})
    return m


def render(model, rays, n_samples, is_train, white_bg, jitter_seed=None):
    """TensorBase.forward of the reference (models/tensorBase.py:436-507).  Training renders draw their per-ray jitter
    from the CPU generator (tensorBase.py:313-317): seed it to replay a known jitter (torch.rand(R, 1) after
    torch.manual_seed(jitter_seed))."""
    if jitter_seed is not None:
        torch.manual_seed(jitter_seed)
    return model(rays, is_train=is_train, white_bg=white_bg, ndc_ray=0, N_samples=n_samples)
