"""Host-side mirror of the reference's ``models/tensoRF.py``: TensorVMSplit
(models/tensoRF.py:139-303) on top of the kernel-backed TensorBase.

Parameters keep the reference's names and shapes ([1,C,H,W] planes, [1,C,L,1] lines,
``basis_mat``, ``renderModule.mlp.*``) so checkpoints and optimisers are interchangeable, but
the plane/line tensors are stored ``channels_last`` -- physically [H][W][C] -- which is the
texel-major layout the gather kernels read with 16-byte vector loads.  Shapes, state-dict
keys and Adam semantics are unchanged ("layout trick").

TensorVM and TensorCP are exported as names only (renderer.py imports them): the reference's
TensorVM is dead code (calls a missing method, tensoRF.py:135) and TensorCP is outside the
north-star path; instantiating either raises.
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.nn.functional as F

from . import _native as nat
from .tensorBase import (AlphaGridMask, TensorBase, _texel_major, positional_encoding, raw2alpha)  # noqa: F401

_CL = torch.channels_last


class _TVFn(torch.autograd.Function):
    """sum_i TVLoss(w)(plane_i) * 1e-2 over the three planes of one factor family through t2n_tv_plane_sums /
    t2n_tv_plane_grad (csrc/tv.cuh).  The gradient pass accumulates straight into the model's gradient buffers: the flat
    all-reduce buffer when enable_flat_grads() is on (autograd then gets nothing, like the render backward)."""

    @staticmethod
    def forward(ctx, model, weight, first_index, track_grad, *planes):
        lib = nat.load()
        dev = planes[0].device
        blocks = int(lib.t2n_tv_blocks())
        cl = [_texel_major(p.detach()) for p in planes]
        partials = torch.empty((len(cl), blocks, 2), device=dev, dtype=torch.float32)
        scales = []
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            for i, t in enumerate(cl):
                _, C, H, W = t.shape
                nat.check(lib.t2n_tv_plane_sums(t.data_ptr(), H, W, C, partials[i].data_ptr(), stream), "t2n_tv_plane_sums")
                # TVLoss: weight * 2 * (h_tv / count_h + w_tv / count_w) / batch, times the 1e-2 of TV_loss_*
                scales.append([weight * 2.0 * 1e-2 / (C * (H - 1) * W), weight * 2.0 * 1e-2 / (C * H * (W - 1))])
        cache = model.__dict__.setdefault("_tv_scale_cache", {})
        key = (first_index, weight, tuple(tuple(t.shape) for t in cl), str(dev))
        if key not in cache:            # device copy of the per-plane normalisations, made once per configuration
            cache[key] = torch.tensor(scales, dtype=torch.float32).to(dev)
        loss = (partials.sum(1) * cache[key]).sum()
        if track_grad and any(ctx.needs_input_grad[4:]):
            ctx.model, ctx.first_index, ctx.scales = model, first_index, scales
            ctx.save_for_backward(*cl)
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        lib = nat.load()
        cl = ctx.saved_tensors
        model = ctx.model
        dev = cl[0].device
        flat = getattr(model, "_flat_grad", None)
        if flat is not None:
            dests = flat["views"][ctx.first_index:ctx.first_index + len(cl)]
        else:
            dests = [torch.zeros_like(t) for t in cl]
        g = g_loss.detach().to(torch.float32).reshape(1).contiguous()
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            for t, d, (ch, cw) in zip(cl, dests, ctx.scales):
                _, C, H, W = t.shape
                nat.check(lib.t2n_tv_plane_grad(t.data_ptr(), H, W, C, g.data_ptr(), ch, cw, d.data_ptr(), stream),
                          "t2n_tv_plane_grad")
        return (None, None, None, None) + ((None,) * len(cl) if flat is not None else tuple(dests))


def _as_list3(n) -> List[int]:
    return [int(n)] * 3 if isinstance(n, int) else [int(v) for v in n]


class _NotBuilt(torch.nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError(f"{type(self).__name__} is outside the B200 hot-path scope (SURVEY.md 2.1 row 2); "
                                  "use TensorVMSplit")


class TensorVM(_NotBuilt):
    pass


class TensorCP(_NotBuilt):
    pass


class TensorVMSplit(TensorBase):
    def __init__(self, aabb, gridSize, device, **kargs):
        super().__init__(aabb, gridSize, device, **kargs)

    # ---- parameters --------------------------------------------------------------------------------
    def init_svd_volume(self, res, device):
        self.density_n_comp = _as_list3(self.density_n_comp)
        self.app_n_comp = _as_list3(self.app_n_comp)
        self.density_plane, self.density_line = self.init_one_svd(self.density_n_comp, self.gridSize, 0.1, device)
        self.app_plane, self.app_line = self.init_one_svd(self.app_n_comp, self.gridSize, 0.1, device)
        self.basis_mat = torch.nn.Linear(sum(self.app_n_comp), self.app_dim, bias=False).to(device)

    def init_one_svd(self, n_component, gridSize, scale, device):
        """0.1*randn factors drawn in the reference's order (plane_i, line_i for i=0,1,2 -- so the
        same torch seed gives the same field, tensoRF.py:150-160), stored channels-last."""
        planes, lines = [], []
        for i in range(3):
            a0, a1 = self.matMode[i]
            v = self.vecMode[i]
            plane = scale * torch.randn((1, n_component[i], int(gridSize[a1]), int(gridSize[a0])))
            line = scale * torch.randn((1, n_component[i], int(gridSize[v]), 1))
            planes.append(torch.nn.Parameter(plane.to(device).contiguous(memory_format=_CL)))
            lines.append(torch.nn.Parameter(line.to(device).contiguous(memory_format=_CL)))
        return torch.nn.ParameterList(planes), torch.nn.ParameterList(lines)

    def get_optparam_groups(self, lr_init_spatialxyz=0.02, lr_init_network=0.001):
        groups = [{'params': self.density_line, 'lr': lr_init_spatialxyz},
                  {'params': self.density_plane, 'lr': lr_init_spatialxyz},
                  {'params': self.app_line, 'lr': lr_init_spatialxyz},
                  {'params': self.app_plane, 'lr': lr_init_spatialxyz},
                  {'params': self.basis_mat.parameters(), 'lr': lr_init_network}]
        if isinstance(self.renderModule, torch.nn.Module):
            groups.append({'params': self.renderModule.parameters(), 'lr': lr_init_network})
        return groups

    # ---- regularisers (parameter-only, dense streaming work; tensoRF.py:173-203) -------------------
    def vectorDiffs(self, vector_comps):
        total = 0
        for comp in vector_comps:
            n_comp, n_size = comp.shape[1:-1]
            flat = comp.reshape(n_comp, n_size)
            gram = flat @ flat.t()
            off_diag = gram.reshape(-1)[1:].view(n_comp - 1, n_comp + 1)[..., :-1]
            total = total + off_diag.abs().mean()
        return total

    def vector_comp_diffs(self):
        return self.vectorDiffs(self.density_line) + self.vectorDiffs(self.app_line)

    def density_L1(self):
        total = 0
        for plane, line in zip(self.density_plane, self.density_line):
            total = total + plane.abs().mean() + line.abs().mean()
        return total

    def TV_loss_density(self, reg):
        """sum_i reg(density_plane[i]) * 1e-2 (tensoRF.py:193-197).  With the reference's utils.TVLoss as `reg` and the
        planes on a CUDA device this runs the fused TV kernels (one read pass for the value, one read + accumulate pass
        for the gradient); any other callable is applied to the planes as tensor ops."""
        return self._tv_loss(reg, list(self.density_plane), 0)

    def TV_loss_app(self, reg):
        """sum_i reg(app_plane[i]) * 1e-2 (tensoRF.py:199-203); see TV_loss_density."""
        return self._tv_loss(reg, list(self.app_plane), 6)

    def _tv_loss(self, reg, planes, first_index):
        w = getattr(reg, "TVLoss_weight", None)
        fused = (w is not None and type(reg).__name__ == "TVLoss" and all(p.is_cuda and p.dtype == torch.float32 for p in planes)
                 and all(p.shape[0] == 1 and p.shape[1] % 4 == 0 and p.shape[2] > 1 and p.shape[3] > 1 for p in planes))
        if not fused:
            total = 0
            for plane in planes:
                total = total + reg(plane) * 1e-2
            return total
        return _TVFn.apply(self, float(w), first_index, torch.is_grad_enabled(), *planes)

    # ---- factor lookups as tensor ops (API parity; the render path gathers inside the kernels) -----
    def _vm_grids(self, xyz_sampled):
        planes = torch.stack([xyz_sampled[..., self.matMode[i]] for i in range(3)]).detach().view(3, -1, 1, 2)
        lines = torch.stack([xyz_sampled[..., self.vecMode[i]] for i in range(3)])
        lines = torch.stack((torch.zeros_like(lines), lines), dim=-1).detach().view(3, -1, 1, 2)
        return planes, lines

    def compute_densityfeature(self, xyz_sampled):
        gp, gl = self._vm_grids(xyz_sampled)
        n = xyz_sampled.shape[0]
        feat = torch.zeros((n,), device=xyz_sampled.device)
        for i in range(3):
            pv = F.grid_sample(self.density_plane[i], gp[[i]], align_corners=True).view(-1, n)
            lv = F.grid_sample(self.density_line[i], gl[[i]], align_corners=True).view(-1, n)
            feat = feat + (pv * lv).sum(0)
        return feat

    def compute_appfeature(self, xyz_sampled):
        gp, gl = self._vm_grids(xyz_sampled)
        n = xyz_sampled.shape[0]
        pv = torch.cat([F.grid_sample(self.app_plane[i], gp[[i]], align_corners=True).view(-1, n) for i in range(3)])
        lv = torch.cat([F.grid_sample(self.app_line[i], gl[[i]], align_corners=True).view(-1, n) for i in range(3)])
        return self.basis_mat((pv * lv).T)

    # ---- grid maintenance (tensoRF.py:243-303) -----------------------------------------------------
    # CUDA-resident factors are resampled / cropped by kernels that read and write the texel-major layout directly
    # (csrc/maint.cu: no NCHW round trip, one launch per factor).  Factors that live on the host (checkpoint surgery,
    # host-logic tests) are reshaped with tensor ops; which one runs is decided by where the data is, never by whether
    # the library happens to be loadable.
    @staticmethod
    def _resample(t, H2, W2):
        """F.interpolate(t, size=(H2, W2), mode='bilinear', align_corners=True) of a [1,C,H,W] factor, channels-last."""
        if not t.is_cuda:
            return F.interpolate(t, size=(H2, W2), mode='bilinear', align_corners=True).contiguous(memory_format=_CL)
        lib = nat.load()
        src = _texel_major(t.detach())
        _, C_, H, W = src.shape
        dst = torch.empty((1, C_, H2, W2), device=t.device, dtype=torch.float32).contiguous(memory_format=_CL)
        with torch.cuda.device(t.device):
            rc = lib.t2n_resample_plane(src.data_ptr(), H, W, C_, dst.data_ptr(), H2, W2,
                                        torch.cuda.current_stream(t.device).cuda_stream)
        nat.check(rc, "t2n_resample_plane")
        return dst

    @staticmethod
    def _crop(t, y0, y1, x0, x1):
        """t[..., y0:y1, x0:x1] of a [1,C,H,W] factor as a new channels-last tensor."""
        if not t.is_cuda:
            return t[..., y0:y1, x0:x1].contiguous(memory_format=_CL)
        lib = nat.load()
        src = _texel_major(t.detach())
        _, C_, H, W = src.shape
        H2, W2 = y1 - y0, x1 - x0
        dst = torch.empty((1, C_, H2, W2), device=t.device, dtype=torch.float32).contiguous(memory_format=_CL)
        with torch.cuda.device(t.device):
            rc = lib.t2n_crop_plane(src.data_ptr(), H, W, C_, y0, x0, dst.data_ptr(), H2, W2,
                                    torch.cuda.current_stream(t.device).cuda_stream)
        nat.check(rc, "t2n_crop_plane")
        return dst

    @torch.no_grad()
    def up_sampling_VM(self, plane_coef, line_coef, res_target):
        for i in range(3):
            a0, a1 = self.matMode[i]
            v = self.vecMode[i]
            plane_coef[i] = torch.nn.Parameter(self._resample(plane_coef[i].data, int(res_target[a1]), int(res_target[a0])))
            line_coef[i] = torch.nn.Parameter(self._resample(line_coef[i].data, int(res_target[v]), 1))
        return plane_coef, line_coef

    @torch.no_grad()
    def upsample_volume_grid(self, res_target):
        self.app_plane, self.app_line = self.up_sampling_VM(self.app_plane, self.app_line, res_target)
        self.density_plane, self.density_line = self.up_sampling_VM(self.density_plane, self.density_line, res_target)
        self.update_stepSize(res_target)
        self._refresh_flat_grads()
        print(f'upsamping to {res_target}')

    @torch.no_grad()
    def shrink(self, new_aabb):
        print("====> shrinking ...")
        xyz_min, xyz_max = new_aabb
        lo = (xyz_min - self.aabb[0]) / self.units
        hi = (xyz_max - self.aabb[0]) / self.units
        lo, hi = torch.round(torch.round(lo)).long(), torch.round(hi).long() + 1
        hi = torch.stack([hi, self.gridSize]).amin(0)
        lo_h, hi_h = [int(v) for v in lo.tolist()], [int(v) for v in hi.tolist()]     # one host read of the six bounds
        for i in range(3):
            v = self.vecMode[i]
            a0, a1 = self.matMode[i]
            for lines in (self.density_line, self.app_line):
                lines[i] = torch.nn.Parameter(self._crop(lines[i].data, lo_h[v], hi_h[v], 0, 1))
            for planes in (self.density_plane, self.app_plane):
                planes[i] = torch.nn.Parameter(self._crop(planes[i].data, lo_h[a1], hi_h[a1], lo_h[a0], hi_h[a0]))
        if not torch.all(self.alphaMask.gridSize == self.gridSize):
            t_lo, t_hi = lo / (self.gridSize - 1), (hi - 1) / (self.gridSize - 1)
            fixed = torch.zeros_like(new_aabb)
            fixed[0] = (1 - t_lo) * self.aabb[0] + t_lo * self.aabb[1]
            fixed[1] = (1 - t_hi) * self.aabb[0] + t_hi * self.aabb[1]
            print("aabb", new_aabb, "\ncorrect aabb", fixed)
            new_aabb = fixed
        self.aabb = new_aabb
        self.update_stepSize((hi_h[0] - lo_h[0], hi_h[1] - lo_h[1], hi_h[2] - lo_h[2]))
        self._refresh_flat_grads()

    # ---- native glue -------------------------------------------------------------------------------
    def _flat_params(self) -> List[torch.Tensor]:
        """Parameter order of the C ABI structs: sigma planes, sigma lines, app planes, app lines,
        basis, then the decoder's six tensors."""
        ps = list(self.density_plane) + list(self.density_line) + list(self.app_plane) + list(self.app_line)
        ps.append(self.basis_mat.weight)
        if self._is_mlp:
            m = self.renderModule.mlp
            ps += [m[0].weight, m[0].bias, m[2].weight, m[2].bias, m[4].weight, m[4].bias]
        return ps

    def _native_param_tensors(self, params: Sequence[torch.Tensor]) -> List[torch.Tensor]:
        out = []
        for j, p in enumerate(params):
            p = p.detach()
            if p.dtype != torch.float32:
                raise TypeError("text2nerf_b200 parameters must be float32")
            out.append(_texel_major(p) if j < 12 else p.contiguous())
        return out

    def _native_params(self, p_cl: Sequence[torch.Tensor]) -> nat.T2NParams:
        s = nat.T2NParams()
        s.sigma_plane = nat.P3(*[t.data_ptr() for t in p_cl[0:3]])
        s.sigma_line = nat.P3(*[t.data_ptr() for t in p_cl[3:6]])
        s.app_plane = nat.P3(*[t.data_ptr() for t in p_cl[6:9]])
        s.app_line = nat.P3(*[t.data_ptr() for t in p_cl[9:12]])
        s.basis = p_cl[12].data_ptr()
        if self._is_mlp:
            s.w1, s.b1, s.w2, s.b2, s.w3, s.b3 = [t.data_ptr() for t in p_cl[13:19]]
            perm, pairs = self._recipe_tensors(p_cl[12].device)
            s.col_perm, s.pair_desc = perm.data_ptr(), pairs.data_ptr()
        return s

    def _grad_buffers(self, p_cl: Sequence[torch.Tensor]) -> List[torch.Tensor]:
        """Zero-filled gradient buffers in the parameters' own memory layout.  With
        enable_flat_grads() they are views into one flat fp32 buffer (the NCCL all-reduce buffer)."""
        flat = getattr(self, "_flat_grad", None)
        if flat is not None:
            # the kernels scatter with the field's CURRENT grid sizes into raw pointers: a view that no longer matches
            # its parameter (factors replaced by upsample / shrink / load without re-enabling) would be written out of
            # bounds, so refuse it here
            for v, t in zip(flat["views"], p_cl):
                if v.shape != t.shape or v.stride() != t.stride():
                    raise RuntimeError("flat gradient buffer is stale: a parameter changed shape/layout "
                                       f"({tuple(v.shape)} vs {tuple(t.shape)}); call enable_flat_grads() again")
            return flat["views"]            # the trainer zeroes the flat buffer once per step
        return [torch.zeros_like(t) for t in p_cl]

    def _native_grads(self, grads: Sequence[torch.Tensor]) -> nat.T2NGrads:
        g = nat.T2NGrads()
        g.sigma_plane = nat.P3(*[t.data_ptr() for t in grads[0:3]])
        g.sigma_line = nat.P3(*[t.data_ptr() for t in grads[3:6]])
        g.app_plane = nat.P3(*[t.data_ptr() for t in grads[6:9]])
        g.app_line = nat.P3(*[t.data_ptr() for t in grads[9:12]])
        g.basis = grads[12].data_ptr()
        if self._is_mlp:
            g.w1, g.b1, g.w2, g.b2, g.w3, g.b3 = [t.data_ptr() for t in grads[13:19]]
        ev = getattr(self, "_app_done_event", None)     # set by text2nerf_b200.dist for the overlapped all-reduce
        g.app_done_event = ev.cuda_event if ev is not None else None
        return g

    def _grads_for_autograd(self, grads: Sequence[torch.Tensor]):
        if getattr(self, "_flat_grad", None) is not None:
            # flat mode: the data gradients live in the flat buffer (all-reduced by the trainer and
            # attached to .grad afterwards, text2nerf_b200/dist.py); autograd gets nothing
            return (None,) * len(grads)
        return tuple(grads)

    def enable_flat_grads(self, on: bool = True):
        """Back every data gradient by ONE flat fp32 buffer (planes, lines, basis, decoder), laid
        out in the C-ABI parameter order.  The ray-sharded trainer all-reduces this buffer with a
        single NCCL call per step (text2nerf_b200/dist.py).  While enabled, backward ACCUMULATES
        into the buffer and hands autograd no parameter gradients: the caller zeroes the buffer
        at the start of a step and attaches the views to .grad after the all-reduce."""
        if not on:
            self._flat_grad = None
            return None
        params = self._flat_params()
        p_cl = self._native_param_tensors(params)
        sizes = [t.numel() for t in p_cl]
        # Buffer order: the appearance factors (C-ABI parameters 6..11) come FIRST -- they are complete earliest in the
        # backward (T2NGrads::app_done_event) and form the segment whose all-reduce overlaps the rest -- then the density
        # factors, basis and decoder.  Every segment stays 16-byte aligned for the vector atomics.
        order = list(range(6, 12)) + list(range(0, 6)) + list(range(12, len(p_cl)))
        offs, o = [0] * len(p_cl), 0
        n_early = 0
        for j in order:
            offs[j] = o
            o += (sizes[j] + 3) & ~3
            if j == 11:
                n_early = o
        buf = torch.zeros((o,), device=p_cl[0].device, dtype=torch.float32)
        views = []
        for t, off in zip(p_cl, offs):
            views.append(torch.as_strided(buf, t.shape, t.stride(), off))
        self._flat_grad = {"buffer": buf, "views": views, "n_early": n_early}
        return buf

    def _refresh_flat_grads(self):
        """The factor Parameters were replaced (upsample_volume_grid / shrink): rebuild the flat gradient buffer and its
        views for the new sizes.  Holders of the old buffer must fetch the new one (enable_flat_grads returns it, and
        model._flat_grad["buffer"] is what dist.allreduce_flat_grads reads)."""
        if getattr(self, "_flat_grad", None) is not None:
            for p in self._flat_params():
                p.grad = None
            self.enable_flat_grads(True)

    def app_sample_count(self) -> int:
        """Samples that passed the weight threshold / the validity test in the last forward
        (device->host read; diagnostics and roofline accounting only)."""
        c = self._last_counters
        return (0, 0) if c is None else tuple(int(v) for v in c[:2].tolist())
