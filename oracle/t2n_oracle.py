"""CPU oracle for the TensoRF ray-marching hot path of eckertzhang/Text2NeRF.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it.  The product path (``text2nerf_b200``) never
does, and fails loudly when its CUDA library is missing.

What it is: a functional restatement (plain functions over a parameter dict,
no nn.Module) of the reference algorithm, executed with PyTorch CPU fp32 ops —
the same third-party arithmetic library (ATen) the reference itself runs on
(pinned torch==1.13.1+cu116 in the reference's requirements.txt:164; the image
has 2.11.0).  Each function cites the reference file:line it follows
(paths relative to the reference checkout).

Pinning: the reference ships NO tests or golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned by running the reference's own
modules in the authoring container: ``tests/golden/make_golden.py`` imports
``models.tensoRF.TensorVMSplit`` from the read-only reference checkout,
renders seeded cases on CPU and writes ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks this file against those fixtures
(bit-exact for the forward outputs) everywhere, and
``tests/test_oracle_live.py`` re-checks against the live reference whenever the
checkout is present.

All tensors are torch CPU tensors; dtype follows the parameters (fp32 for
parity, fp64 only for the noise-floor study in tests).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

# plane i spans axes MAT_AXES[i] = (width axis, height axis); line i runs along VEC_AXIS[i]
# models/tensorBase.py:190-191
MAT_AXES = ((0, 1), (0, 2), (1, 2))
VEC_AXIS = (2, 1, 0)

SHADING_MODES = ("MLP_Fea_noview", "MLP_Fea", "MLP_PE", "MLP", "SH", "RGB")


@dataclass
class FieldSpec:
    """Scalar configuration of one TensorVMSplit field (ctor kwargs of
    models/tensorBase.py:164-198 plus the grid)."""
    aabb: Sequence[Sequence[float]]           # [[x0,y0,z0],[x1,y1,z1]]
    grid: Sequence[int]                       # [Gx, Gy, Gz]
    near_far: Sequence[float] = (0.5, 8.0)
    step_ratio: float = 1.0
    density_shift: float = -10.0
    distance_scale: float = 25.0
    weight_thres: float = 1e-4                # rayMarch_weight_thres
    act: str = "softplus"                     # fea2denseAct
    shading: str = "MLP_Fea_noview"
    pos_pe: int = 6
    view_pe: int = 2
    fea_pe: int = 6
    featureC: int = 128
    app_dim: int = 27
    density_n_comp: Sequence[int] = (16, 16, 16)
    app_n_comp: Sequence[int] = (48, 48, 48)
    eval_z_min: float = 2.0                   # hard-coded world-z filter, tensorBase.py:459-462
    dtype: torch.dtype = torch.float32

    def aabb_t(self) -> torch.Tensor:
        return torch.tensor(self.aabb, dtype=self.dtype)


def derive_step(spec: FieldSpec):
    """stepSize (0-dim tensor) and nSamples exactly as update_stepSize does
    (models/tensorBase.py:220-231): all tensor arithmetic in the field dtype."""
    aabb = spec.aabb_t()
    size = aabb[1] - aabb[0]
    grid = torch.tensor(list(spec.grid), dtype=torch.long)
    units = size / (grid - 1)
    step = torch.mean(units) * spec.step_ratio
    diag = torch.sqrt(torch.sum(torch.square(size)))
    n_samples = int((diag / step).item()) + 1
    return step, n_samples


def mlp_in_dim(spec: FieldSpec) -> int:
    """Input width of the appearance decoder (models/tensorBase.py:66,92,115,141)."""
    c = spec.app_dim
    if spec.shading == "MLP_Fea_noview":
        return 2 * spec.fea_pe * c + c
    if spec.shading == "MLP_Fea":
        return 2 * spec.view_pe * 3 + 2 * spec.fea_pe * c + 3 + c
    if spec.shading == "MLP_PE":
        return (3 + 2 * spec.view_pe * 3) + (3 + 2 * spec.pos_pe * 3) + c
    if spec.shading == "MLP":
        return (3 + 2 * spec.view_pe * 3) + c
    return 0


def init_params(spec: FieldSpec, seed: int = 0, density_gain: float = 1.0,
                app_gain: float = 1.0) -> Dict[str, torch.Tensor]:
    """Seeded synthetic parameters with the reference's state-dict keys and
    shapes (models/tensoRF.py:144-160, tensorBase.py:94-99).  Not an RNG-exact
    replay of the reference initialiser; tests that need the reference's own
    init load its state_dict instead."""
    g = torch.Generator().manual_seed(seed)
    G = list(spec.grid)
    p: Dict[str, torch.Tensor] = {}

    def randn(*shape):
        return torch.randn(*shape, generator=g, dtype=torch.float32).to(spec.dtype)

    for i in range(3):
        a0, a1 = MAT_AXES[i]
        v = VEC_AXIS[i]
        p[f"density_plane.{i}"] = 0.1 * density_gain * randn(1, spec.density_n_comp[i], G[a1], G[a0])
        p[f"density_line.{i}"] = 0.1 * density_gain * randn(1, spec.density_n_comp[i], G[v], 1)
        p[f"app_plane.{i}"] = 0.1 * app_gain * randn(1, spec.app_n_comp[i], G[a1], G[a0])
        p[f"app_line.{i}"] = 0.1 * app_gain * randn(1, spec.app_n_comp[i], G[v], 1)
    n_app = sum(spec.app_n_comp)
    p["basis_mat.weight"] = randn(spec.app_dim, n_app) / math.sqrt(n_app)
    k = mlp_in_dim(spec)
    if k > 0:
        C = spec.featureC
        p["renderModule.mlp.0.weight"] = randn(C, k) / math.sqrt(k)
        p["renderModule.mlp.0.bias"] = 0.1 * randn(C)
        p["renderModule.mlp.2.weight"] = randn(C, C) / math.sqrt(C)
        p["renderModule.mlp.2.bias"] = 0.1 * randn(C)
        p["renderModule.mlp.4.weight"] = randn(3, C) / math.sqrt(C)
        p["renderModule.mlp.4.bias"] = torch.zeros(3, dtype=spec.dtype)
    return p


# ----------------------------------------------------------------------------------------------
# sampling
# ----------------------------------------------------------------------------------------------

def march_points(spec: FieldSpec, rays_o, rays_d, n_samples: int, jitter: Optional[torch.Tensor]):
    """Sample generation of TensorBase.sample_ray (models/tensorBase.py:304-323).

    ``jitter`` is the per-ray U[0,1) offset the reference draws from the CPU RNG
    when is_train (tensorBase.py:313-316), shape [R] or [R,1]; None = eval."""
    step, _ = derive_step(spec)
    aabb = spec.aabb_t()
    near, far = spec.near_far
    vec = torch.where(rays_d == 0, torch.full_like(rays_d, 1e-6), rays_d)
    rate_hi = (aabb[1] - rays_o) / vec
    rate_lo = (aabb[0] - rays_o) / vec
    t_min = torch.minimum(rate_hi, rate_lo).amax(-1).clamp(min=near, max=far)
    idx = torch.arange(n_samples)[None].float()
    if jitter is not None:
        idx = idx.repeat(rays_d.shape[-2], 1)
        idx += jitter.reshape(-1, 1).float()
    idx = idx.to(spec.dtype)
    z = t_min[..., None] + step * idx
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z[..., None]
    outside = ((aabb[0] > pts) | (pts > aabb[1])).any(dim=-1)
    return pts, z, ~outside


def to_unit_cube(spec: FieldSpec, pts):
    """normalize_coord (models/tensorBase.py:245-246) with invaabbSize of :224."""
    aabb = spec.aabb_t()
    inv = 2.0 / (aabb[1] - aabb[0])
    return (pts - aabb[0]) * inv - 1


# ----------------------------------------------------------------------------------------------
# VM factor lookups
# ----------------------------------------------------------------------------------------------

def _vm_coords(xn):
    """Coordinate packing of compute_densityfeature/compute_appfeature
    (models/tensoRF.py:208-210, 226-228): plane grids (a0,a1), line grids (0,v)."""
    cp = torch.stack([xn[..., list(MAT_AXES[i])] for i in range(3)]).detach().view(3, -1, 1, 2)
    cl = torch.stack([xn[..., VEC_AXIS[i]] for i in range(3)])
    cl = torch.stack((torch.zeros_like(cl), cl), dim=-1).detach().view(3, -1, 1, 2)
    return cp, cl


def density_feature(params, xn):
    """sigma feature = sum_i sum_c plane_i[c](x) * line_i[c](x)
    (models/tensoRF.py:205-220)."""
    cp, cl = _vm_coords(xn)
    n = xn.shape[0]
    out = torch.zeros((n,), dtype=xn.dtype)
    for i in range(3):
        pv = F.grid_sample(params[f"density_plane.{i}"], cp[[i]], align_corners=True).view(-1, n)
        lv = F.grid_sample(params[f"density_line.{i}"], cl[[i]], align_corners=True).view(-1, n)
        out = out + torch.sum(pv * lv, dim=0)
    return out


def app_products(params, xn):
    """[n, sum(app_n_comp)] plane*line products before the basis matrix
    (models/tensoRF.py:223-239 up to the basis_mat call)."""
    cp, cl = _vm_coords(xn)
    n = xn.shape[0]
    pv, lv = [], []
    for i in range(3):
        pv.append(F.grid_sample(params[f"app_plane.{i}"], cp[[i]], align_corners=True).view(-1, n))
        lv.append(F.grid_sample(params[f"app_line.{i}"], cl[[i]], align_corners=True).view(-1, n))
    return (torch.cat(pv) * torch.cat(lv)).T


def app_feature(params, xn):
    """compute_appfeature (models/tensoRF.py:223-239): basis_mat is a bias-free Linear."""
    return F.linear(app_products(params, xn), params["basis_mat.weight"])


def density_activation(spec: FieldSpec, feat):
    """feature2density (models/tensorBase.py:406-410)."""
    if spec.act == "softplus":
        return F.softplus(feat + spec.density_shift)
    if spec.act == "relu":
        return F.relu(feat)
    raise ValueError(spec.act)


# ----------------------------------------------------------------------------------------------
# appearance decoders
# ----------------------------------------------------------------------------------------------

def freq_encode(x, n_freq: int):
    """positional_encoding (models/tensorBase.py:11-17): channel-major, frequency-minor,
    all sines then all cosines."""
    bands = (2 ** torch.arange(n_freq).float()).to(x.dtype)
    y = (x[..., None] * bands).reshape(x.shape[:-1] + (n_freq * x.shape[-1],))
    return torch.cat([torch.sin(y), torch.cos(y)], dim=-1)


def sh_basis_deg2(dirs):
    """eval_sh_bases(2, dirs) (models/sh.py:87-111): 9 real SH basis values."""
    c0 = 0.28209479177387814
    c1 = 0.4886025119029199
    c2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005,
          -1.0925484305920792, 0.5462742152960396)
    x, y, z = dirs.unbind(-1)
    xx, yy, zz = x * x, y * y, z * z
    xy, yz, xz = x * y, y * z, x * z
    out = torch.empty((*dirs.shape[:-1], 9), dtype=dirs.dtype)
    out[..., 0] = c0
    out[..., 1] = -c1 * y
    out[..., 2] = c1 * z
    out[..., 3] = -c1 * x
    out[..., 4] = c2[0] * xy
    out[..., 5] = c2[1] * yz
    out[..., 6] = c2[2] * (2.0 * zz - xx - yy)
    out[..., 7] = c2[3] * xz
    out[..., 8] = c2[4] * (xx - yy)
    return out


def decode_rgb(spec: FieldSpec, params, xn, viewdirs, feat):
    """The six shading heads of init_render_func (models/tensorBase.py:200-218):
    MLPRender_Fea_noview :88-109, MLPRender_Fea :62-86, MLPRender_PE :111-135,
    MLPRender :137-159, SHRender :29-33, RGBRender :36-39."""
    mode = spec.shading
    if mode == "RGB":
        return feat
    if mode == "SH":
        mult = sh_basis_deg2(viewdirs)[:, None]
        coeff = feat.view(-1, 3, mult.shape[-1])
        return torch.relu(torch.sum(mult * coeff, dim=-1) + 0.5)
    if mode == "MLP_Fea_noview":
        cols = [feat]
        if spec.fea_pe > 0:
            cols.append(freq_encode(feat, spec.fea_pe))
    elif mode == "MLP_Fea":
        cols = [feat, viewdirs]
        if spec.fea_pe > 0:
            cols.append(freq_encode(feat, spec.fea_pe))
        if spec.view_pe > 0:
            cols.append(freq_encode(viewdirs, spec.view_pe))
    elif mode == "MLP_PE":
        cols = [feat, viewdirs]
        if spec.pos_pe > 0:
            cols.append(freq_encode(xn, spec.pos_pe))
        if spec.view_pe > 0:
            cols.append(freq_encode(viewdirs, spec.view_pe))
    elif mode == "MLP":
        cols = [feat, viewdirs]
        if spec.view_pe > 0:
            cols.append(freq_encode(viewdirs, spec.view_pe))
    else:
        raise ValueError(mode)
    h = torch.cat(cols, dim=-1)
    h = torch.relu(F.linear(h, params["renderModule.mlp.0.weight"], params["renderModule.mlp.0.bias"]))
    h = torch.relu(F.linear(h, params["renderModule.mlp.2.weight"], params["renderModule.mlp.2.bias"]))
    h = F.linear(h, params["renderModule.mlp.4.weight"], params["renderModule.mlp.4.bias"])
    return torch.sigmoid(h)


def decoder_preactivations(spec: FieldSpec, params, xn, viewdirs, feat):
    """(h1, h2): the two hidden pre-activations of the MLP heads (models/tensorBase.py:94-96) for given samples.
    Used by gradient-parity checks to find samples that sit on a ReLU kink: d relu / dh flips when an implementation's
    h differs from the reference's by more than |h| (fp32 summation order moves h by up to ~2e-6 here), and the
    gradient of such a sample changes by a finite amount -- in the reference as much as in any re-implementation."""
    mode = spec.shading
    cols = [feat] if mode == "MLP_Fea_noview" else [feat, viewdirs]
    if mode in ("MLP_Fea_noview", "MLP_Fea") and spec.fea_pe > 0:
        cols.append(freq_encode(feat, spec.fea_pe))
    if mode in ("MLP_Fea", "MLP") and spec.view_pe > 0:
        cols.append(freq_encode(viewdirs, spec.view_pe))
    h1 = F.linear(torch.cat(cols, dim=-1), params["renderModule.mlp.0.weight"], params["renderModule.mlp.0.bias"])
    h2 = F.linear(torch.relu(h1), params["renderModule.mlp.2.weight"], params["renderModule.mlp.2.bias"])
    return h1, h2


def relu_kink_samples(spec: FieldSpec, params, rays, aux, eps: float = 4e-6):
    """Number of listed samples (aux = render(..., keep=True)[4]) with a hidden unit within eps of a ReLU kink, and
    the largest compositing weight among them."""
    m = aux["app_mask"]
    if spec.shading not in ("MLP_Fea_noview", "MLP_Fea", "MLP") or not bool(m.any()):
        return 0, 0.0
    xn = aux["xn"][m]
    viewdirs = rays[:, 3:6].view(-1, 1, 3).expand(aux["xn"].shape)[m]
    with torch.no_grad():
        p = {k: v.detach() for k, v in params.items()}
        h1, h2 = decoder_preactivations(spec, p, xn, viewdirs, app_feature(p, xn))
        near = (h1.abs().min(dim=1).values < eps) | (h2.abs().min(dim=1).values < eps)
    if not bool(near.any()):
        return 0, 0.0
    w = aux.get("weight_listed")
    return int(near.sum()), float(w[near].max()) if w is not None else float("nan")


# ----------------------------------------------------------------------------------------------
# compositing
# ----------------------------------------------------------------------------------------------

def alpha_composite(sigma, dist):
    """raw2alpha (models/tensorBase.py:19-26)."""
    alpha = 1. - torch.exp(-sigma * dist)
    ones = torch.ones(alpha.shape[0], 1, dtype=alpha.dtype)
    trans = torch.cumprod(torch.cat([ones, 1. - alpha + 1e-10], -1), -1)
    weight = alpha * trans[:, :-1]
    return alpha, weight, trans[:, -1:]


def alpha_mask_lookup(mask_volume, mask_aabb, pts):
    """AlphaGridMask.sample_alpha (models/tensorBase.py:52-59): trilinear lookup of a
    [1,1,Z,Y,X] occupancy volume in its own aabb."""
    inv = 1.0 / (mask_aabb[1] - mask_aabb[0]) * 2
    xn = (pts - mask_aabb[0]) * inv - 1
    return F.grid_sample(mask_volume, xn.view(1, -1, 1, 1, 3), align_corners=True).view(-1)


def render(spec: FieldSpec, params: Dict[str, torch.Tensor], rays, n_samples: int = -1,
           is_train: bool = False, white_bg: bool = True, jitter: Optional[torch.Tensor] = None,
           alpha_mask=None, keep: bool = False, app_mask_override: Optional[torch.Tensor] = None):
    """TensorBase.forward for ndc_ray=False (models/tensorBase.py:436-507).

    ``rays`` [R,6] = origin, direction.  ``jitter`` must be given iff is_train.
    ``alpha_mask`` = (volume[1,1,Z,Y,X], aabb[2,3]) or None.
    ``app_mask_override`` [R,S] bool replaces the selection ``weight > rayMarch_weight_thres`` of tensorBase.py:477: the
    selection is a discontinuity of the reference's function (a weight within fp32 noise of the threshold flips it even
    between the reference in fp32 and in fp64), so gradient comparisons evaluate the reference's function
    on the selection the implementation under test made whenever the two selections differ in isolated samples.
    Returns (rgb_map[R,3], depth_map[R], z_vals[R,S], weight[R,S]) and, with
    keep=True, a dict of intermediates as a fifth element."""
    if is_train and jitter is None:
        raise ValueError("training render needs the per-ray jitter the reference draws on the CPU")
    if not is_train:
        jitter = None
    if n_samples <= 0:
        n_samples = derive_step(spec)[1]
    o, d = rays[:, :3], rays[:, 3:6]
    pts, z, valid = march_points(spec, o, d, n_samples, jitter)
    dists = torch.cat((z[:, 1:] - z[:, :-1], torch.zeros_like(z[:, :1])), dim=-1)
    viewdirs = d.view(-1, 1, 3).expand(pts.shape)

    if alpha_mask is not None:                                   # tensorBase.py:451-456
        vol, maabb = alpha_mask
        occ = alpha_mask_lookup(vol, maabb, pts[valid]) > 0
        invalid = ~valid
        invalid[valid] |= (~occ)
        valid = ~invalid
    if not is_train:                                             # tensorBase.py:459-462
        valid = valid * (pts[:, :, -1] > spec.eval_z_min)

    sigma = torch.zeros(pts.shape[:-1], dtype=pts.dtype)
    rgb = torch.zeros((*pts.shape[:2], 3), dtype=pts.dtype)
    xn = pts
    sigma_feat = None
    if valid.any():                                              # tensorBase.py:467-472
        xn = to_unit_cube(spec, pts)
        sigma_feat = density_feature(params, xn[valid])
        sigma[valid] = density_activation(spec, sigma_feat)

    alpha, weight, bg = alpha_composite(sigma, dists * spec.distance_scale)
    app_mask = weight > spec.weight_thres
    if app_mask_override is not None:
        app_mask = app_mask_override
    if app_mask.any():                                           # tensorBase.py:489-492
        feat = app_feature(params, xn[app_mask])
        rgb[app_mask] = decode_rgb(spec, params, xn[app_mask], viewdirs[app_mask], feat)

    acc = torch.sum(weight, -1)
    rgb_map = torch.sum(weight[..., None] * rgb, -2)
    if white_bg:                                                 # tensorBase.py:497 (RNG branch is the caller's)
        rgb_map = rgb_map + (1. - acc[..., None])
    rgb_map = rgb_map.clamp(0, 1)
    depth_map = torch.sum(weight * z, -1)
    depth_map = depth_map + (1. - acc) * rays[..., -1]           # last ray column = d_z (tensorBase.py:505)
    if keep:
        aux = dict(sigma=sigma, alpha=alpha, valid=valid, app_mask=app_mask, rgb=rgb, acc=acc,
                   bg=bg, xn=xn, sigma_feat=sigma_feat, weight_listed=weight.detach()[app_mask])
        return rgb_map, depth_map, z, weight, aux
    return rgb_map, depth_map, z, weight


# ----------------------------------------------------------------------------------------------
# callers either side of the path
# ----------------------------------------------------------------------------------------------

def render_chunked(spec, params, rays, chunk=4096, **kw):
    """OctreeRender_trilinear_fast (renderer.py:28-42): chunk loop + concatenation.
    Returns (rgb, None, depth, weights, z_vals) like the reference."""
    outs: List[List[torch.Tensor]] = [[], [], [], []]
    jit = kw.pop("jitter", None)
    n = rays.shape[0]
    for c in range(n // chunk + int(n % chunk > 0)):
        sl = slice(c * chunk, (c + 1) * chunk)
        r = render(spec, params, rays[sl], jitter=None if jit is None else jit[sl], **kw)
        for lst, t in zip(outs, r):
            lst.append(t)
    rgb, depth, z, w = (torch.cat(x) for x in outs)
    return rgb, None, depth, w, z


def pixel_directions(H: int, W: int, focal, center=None):
    """get_ray_directions (dataLoader/ray_utils.py:24-42): pixel centres at +0.5,
    OpenCV axes, z = 1 (callers normalise, dataLoader/scene_gen.py:45)."""
    fx, fy = (focal, focal) if not isinstance(focal, (list, tuple)) else focal
    xs = torch.linspace(0, W - 1, W)
    ys = torch.linspace(0, H - 1, H)
    j, i = torch.meshgrid(ys, xs, indexing="ij")
    i = i + 0.5
    j = j + 0.5
    cx, cy = center if center is not None else (W / 2, H / 2)
    return torch.stack([(i - cx) / fx, (j - cy) / fy, torch.ones_like(i)], -1)


def camera_rays(directions, c2w):
    """get_rays (dataLoader/ray_utils.py:66-87): rotate, broadcast origin, no renormalisation."""
    rd = directions @ c2w[:3, :3].T
    ro = c2w[:3, 3].expand(rd.shape)
    return ro.reshape(-1, 3), rd.reshape(-1, 3)


def training_loss(rgb_map, depth_map, z_vals, weight, rgb_gt, depth_gt,
                  w_depth: float = 0.005, w_trans: float = 1e3, delta: float = 0.1):
    """Data terms of the Text2NeRF training step (text2nerf_main.py:563-575) with
    TransMittanceLoss_mask (utils.py:67-80)."""
    l_rgb = torch.mean((rgb_map - rgb_gt) ** 2)
    l_depth = torch.mean((depth_map - depth_gt) ** 2)
    mask = (z_vals - depth_gt[:, None] + delta) < 0
    mean_w = torch.mean(weight * mask, dim=1)
    l_trans = torch.mean((mean_w - torch.zeros_like(mean_w)) ** 2)
    return l_rgb + w_depth * l_depth + w_trans * l_trans


def tv_plane(x):
    """TVLoss(1) (utils.py:488-504) on one [1,C,H,W] plane."""
    n_h = x[:, :, 1:, :].numel()
    n_w = x[:, :, :, 1:].numel()
    h_tv = torch.pow(x[:, :, 1:, :] - x[:, :, :-1, :], 2).sum()
    w_tv = torch.pow(x[:, :, :, 1:] - x[:, :, :, :-1], 2).sum()
    return 2 * (h_tv / n_h + w_tv / n_w) / x.shape[0]


# ----------------------------------------------------------------------------------------------
# index-form restatement of the bilinear lookup (documents the rounding sequence the CUDA
# kernels must reproduce; checked against F.grid_sample in tests)
# ----------------------------------------------------------------------------------------------

def texel_coords(xn_axis, size: int):
    """align_corners=True un-normalisation ((x+1)/2)*(size-1) of ATen
    (torch/include/ATen/native/GridSampler.h:27-36), floor and fraction."""
    ix = ((xn_axis + 1) / 2) * (size - 1)
    i0 = torch.floor(ix)
    return i0.long(), ix - i0


def density_feature_indexform(params, xn, grid):
    """Same value as density_feature up to summation order, written as explicit taps:
    4 plane texels (weights (1-fx)(1-fy), fx(1-fy), (1-fx)fy, fx fy; per-corner bounds
    check = zeros padding) and 2 line texels per factor."""
    n = xn.shape[0]
    out = torch.zeros((n,), dtype=xn.dtype)
    for i in range(3):
        a0, a1 = MAT_AXES[i]
        v = VEC_AXIS[i]
        P = params[f"density_plane.{i}"][0]            # [C,H,W]
        L = params[f"density_line.{i}"][0, :, :, 0]    # [C,L]
        C, H, W = P.shape
        x0, fx = texel_coords(xn[:, a0], W)
        y0, fy = texel_coords(xn[:, a1], H)
        l0, fl = texel_coords(xn[:, v], L.shape[1])
        pv = torch.zeros((C, n), dtype=xn.dtype)
        for dy, wy in ((0, 1 - fy), (1, fy)):
            for dx, wx in ((0, 1 - fx), (1, fx)):
                xi, yi = x0 + dx, y0 + dy
                ok = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
                tap = P[:, yi.clamp(0, H - 1), xi.clamp(0, W - 1)]
                pv = pv + torch.where(ok, wx * wy, torch.zeros_like(wx)) * tap
        lv = torch.zeros((C, n), dtype=xn.dtype)
        for dl, wl in ((0, 1 - fl), (1, fl)):
            li = l0 + dl
            ok = (li >= 0) & (li < L.shape[1])
            lv = lv + torch.where(ok, wl, torch.zeros_like(wl)) * L[:, li.clamp(0, L.shape[1] - 1)]
        out = out + torch.sum(pv * lv, dim=0)
    return out
