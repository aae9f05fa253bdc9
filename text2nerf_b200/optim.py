"""Fused multi-tensor Adam for the training loop.

``FusedAdam`` is a drop-in for ``torch.optim.Adam(grad_vars, betas=(0.9, 0.99))`` as text2nerf_main.py:453-454 builds it
from ``tensorf.get_optparam_groups(lr_xyz, lr_net)``: same constructor arguments, same ``param_groups`` (the loop decays
``param_group['lr']`` in place every iteration, text2nerf_main.py:597-598), same state keys (``step``, ``exp_avg``,
``exp_avg_sq``) so optimiser state dicts are interchangeable.  ``step()`` is ONE kernel launch over all parameter
tensors (csrc/adam.cuh) instead of torch's per-op multi-tensor passes.  CUDA float32 parameters only; anything else
raises -- there is no silent fallback."""
import ctypes as C

import torch

from . import _native as nat


def _dense(t: torch.Tensor) -> bool:
    return t.is_contiguous() or (t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last))


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = nat.load()
        # one launch per distinct (betas, eps, weight_decay, step) combination; the Text2NeRF groups differ in lr only
        buckets = {}
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32:
                    raise nat.NativeLibraryError("FusedAdam handles CUDA float32 parameters only")
                if p.grad.is_sparse:
                    raise RuntimeError("FusedAdam does not support sparse gradients")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] = int(st["step"]) + 1
                g = p.grad
                if not _dense(p):
                    raise RuntimeError("FusedAdam needs dense parameters (contiguous or channels_last)")
                if g.stride() != p.stride():
                    # bring the gradient into the parameter's layout (rare: the kernels already write it there)
                    g = torch.empty_like(p, memory_format=torch.preserve_format).copy_(g)
                for name in ("exp_avg", "exp_avg_sq"):
                    if st[name].stride() != p.stride():
                        st[name] = torch.empty_like(p, memory_format=torch.preserve_format).copy_(st[name])
                key = (p.device, group["betas"], group["eps"], group["weight_decay"], st["step"])
                buckets.setdefault(key, []).append((p, g, st["exp_avg"], st["exp_avg_sq"], float(group["lr"])))
        for (dev, betas, eps, wd, step), items in buckets.items():
            arr = (nat.T2NAdamTensor * len(items))()
            for i, (p, g, m, v, lr) in enumerate(items):
                arr[i] = nat.T2NAdamTensor(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr)
            with torch.cuda.device(dev):
                rc = lib.t2n_adam_step(arr, len(items), float(betas[0]), float(betas[1]), float(eps), float(wd), int(step),
                                       torch.cuda.current_stream(dev).cuda_stream)
            nat.check(rc, "t2n_adam_step")
        return loss
