"""world_size-2 CPU tests (gloo) of the sharding / flat-gradient host logic.  The kernels are
not involved (no GPU here); what is checked is that slicing + one all-reduce over the flat
buffer reproduces the single-process gradient, with the collective being the only exchange."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import Case, build_model
from text2nerf_b200 import dist as t2n_dist


def test_shard_bounds_partition():
    for n in (1, 7, 64, 640000, 4097):
        for w in (1, 2, 3, 4, 8):
            spans = [t2n_dist.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert t2n_dist.shard_views(10, 1, 4) == [1, 5, 9]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c = Case("t2n_noview_train")
        model = build_model(c.spec, c.params, "cpu")
        flat = model.enable_flat_grads(True)
        flat.zero_()
        # stand-in for the backward kernels: every rank ACCUMULATES the gradient of its ray slice
        # (here: a deterministic function of the slice) into the flat views
        R = 64
        lo, hi = t2n_dist.shard_bounds(R, rank, world)
        g = torch.Generator().manual_seed(0)
        per_ray = [torch.randn(R, *([1] * v.dim()), generator=g) for v in model._flat_grad["views"]]
        for v, pr in zip(model._flat_grad["views"], per_ray):
            v.add_((pr[lo:hi] * torch.ones_like(v)[None]).sum(0))
        t2n_dist.allreduce_flat_grads(model, world)
        # a parameter-only regulariser gradient already sitting in .grad must survive attach
        p0 = model._flat_params()[0]
        p0.grad = torch.full_like(p0, 0.5)
        t2n_dist.attach_flat_grads(model)
        expect = [(pr * 1.0).sum(0).expand_as(v) / world for v, pr in zip(model._flat_grad["views"], per_ray)]
        ok = all(torch.allclose(p.grad if i else p.grad - 0.5, e, atol=1e-5)
                 for i, (p, e) in enumerate(zip(model._flat_params(), expect)))
        # the two-segment form (appearance factors first, 1/world folded into the loss by the caller): plain sums
        flat.zero_()
        for p in model._flat_params():
            p.grad = None
        for v, pr in zip(model._flat_grad["views"], per_ray):
            v.add_((pr[lo:hi] * torch.ones_like(v)[None]).sum(0) / world)
        n_early = model._flat_grad["n_early"]
        assert n_early == sum((v.numel() + 3) & ~3 for v in model._flat_grad["views"][6:12])
        assert model._flat_grad["views"][6].data_ptr() == flat.data_ptr()        # app_plane.0 leads the buffer
        t2n_dist.allreduce_flat_grads_overlapped(model, world)
        ok = ok and all(torch.allclose(v, e, atol=1e-5) for v, e in zip(model._flat_grad["views"], expect))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_flat_allreduce_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]
