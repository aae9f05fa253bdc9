"""Ray/view sharding across the GPUs of one box.

The reference is single-GPU; this is new functionality required by the north star: one process
per GPU, parameters replicated, every rank renders / differentiates its own contiguous slice of
the ray batch, and a training step ends with exactly ONE collective -- an all-reduce (sum) over
the flat fp32 gradient buffer that the backward kernels accumulated into
(TensorVMSplit.enable_flat_grads) -- followed by a division by the world size (all three data
losses are means over rays, text2nerf_main.py:563-575, utils.py:77-78).

Equivalence with the single-GPU step needs (tests/test_dist_gloo.py):
  * the per-ray jitter of the FULL batch is drawn with the same CPU seed on every rank and sliced;
  * parameter-only regularisers (TV) are evaluated on every rank and added after the averaged
    data gradient (attach_flat_grads adds into an existing .grad).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of n items for `rank`; sizes differ by at most one, slices of
    neighbouring ranks are adjacent (contiguous rows keep texel locality within a view)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_rays(rays: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    lo, hi = shard_bounds(rays.shape[0], rank, world)
    return rays[lo:hi]


def shard_views(n_views: int, rank: int, world: int):
    """Round-robin view assignment for rendering (no communication)."""
    return list(range(rank, n_views, world))


def allreduce_flat_grads(model, world: int, group=None) -> torch.Tensor:
    """The single collective of a training step: all-reduce(sum) of the flat gradient buffer over
    NCCL (NVLink/NVSwitch), then scale by 1/world.  No-op scaling for world == 1."""
    flat = model._flat_grad["buffer"]
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.mul_(1.0 / world)
    return flat


def enable_overlapped_allreduce(model, on: bool = True) -> None:
    """Prepare the two-segment all-reduce: the backward records an event when the gradients of the appearance factors are
    complete (T2NGrads::app_done_event, right behind the scatter kernel) -- the weight-gradient GEMMs and the density sweep
    still run on the other streams of the backward's fork -- and the all-reduce of that segment (the front 75 % of the flat
    buffer at 16/48 components) goes out on a communication stream behind the event, overlapping the rest of the
    backward.  The remainder (density factors, basis, decoder) follows when the backward has joined."""
    if not on:
        model._app_done_event = None
        model._comm_stream = None
        return
    dev = model._flat_grad["buffer"].device
    with torch.cuda.device(dev):
        model._app_done_event = torch.cuda.Event()
        model._comm_stream = torch.cuda.Stream(device=dev)
    # make sure the event handle exists before the first backward hands it to the library
    model._app_done_event.record(torch.cuda.current_stream(dev))


def allreduce_flat_grads_overlapped(model, world: int, group=None) -> torch.Tensor:
    """All-reduce(sum) of the flat gradient buffer in two segments, the appearance segment overlapped with the density
    sweep (enable_overlapped_allreduce).  No scaling pass: the caller folds 1/world into the loss
    (data_loss(n_rays_total=world * rays_per_rank)), so the summed gradients are already the full-batch mean."""
    fg = model._flat_grad
    flat, n_early = fg["buffer"], fg["n_early"]
    if world <= 1:
        return flat
    if not flat.is_cuda:                    # host tensors (gloo tests of the segment logic): no streams to overlap on
        dist.all_reduce(flat[:n_early], op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(flat[n_early:], op=dist.ReduceOp.SUM, group=group)
        return flat
    dev = flat.device
    cur = torch.cuda.current_stream(dev)
    comm = model._comm_stream
    comm.wait_event(model._app_done_event)
    with torch.cuda.stream(comm):
        w_app = dist.all_reduce(flat[:n_early], op=dist.ReduceOp.SUM, group=group, async_op=True)
    w_den = dist.all_reduce(flat[n_early:], op=dist.ReduceOp.SUM, group=group, async_op=True)
    w_app.wait()
    w_den.wait()
    cur.wait_stream(comm)
    return flat


def attach_flat_grads(model) -> None:
    """Make the (all-reduced) flat-buffer views the parameters' .grad; gradients that autograd
    already accumulated there (parameter-only regularisers) are kept and added."""
    for p, v in zip(model._flat_params(), model._flat_grad["views"]):
        if p.grad is None:
            p.grad = v
        elif p.grad.data_ptr() != v.data_ptr():
            p.grad.add_(v)


def sharded_train_step(model, optimizer, rays, targets_fn, n_samples, rank: int, world: int, white_bg=True,
                       regulariser=None, seed=None):
    """One data-parallel step.  `rays` is the FULL batch [R,6] (identical on every rank);
    `targets_fn(lo, hi, outputs)` returns this rank's share of the summed-over-rays loss divided
    by the FULL batch size R (so the all-reduced sum of the rank losses is the single-GPU mean)."""
    R = rays.shape[0]
    lo, hi = shard_bounds(R, rank, world)
    flat = model._flat_grad["buffer"]
    flat.zero_()
    optimizer.zero_grad(set_to_none=True)
    if seed is not None:
        torch.manual_seed(seed)
    jitter_full = torch.rand(R, 1)                      # same CPU draw on every rank (tensorBase.py:316)
    # tensorBase.py:497: training composites a white background at random when white_bg is False; the draw follows the
    # jitter draw like in TensorBase.forward and is identical on every rank (same seed)
    white = bool(white_bg) or bool(torch.rand((1,)) < 0.5)
    from .tensorBase import _RenderFn
    out = _RenderFn.apply(model, rays[lo:hi].contiguous(), jitter_full[lo:hi].to(rays.device).view(-1).contiguous(),
                          n_samples, True, white, True, *model._flat_params())
    loss = targets_fn(lo, hi, out) * world               # undo the 1/world of the all-reduce average
    loss.backward()
    allreduce_flat_grads(model, world)
    if regulariser is not None:
        regulariser(model).backward()                    # parameter-only terms: identical on every rank
    attach_flat_grads(model)
    optimizer.step()
    return loss.detach() / world
