"""Host-side multi-rank logic on CPU with the gloo backend, world_size 2:
naiveSyncBN1d statistics exchange and the flat-gradient all-reduce convention of FlatTrainer."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import geomae_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from geomae_b200.norm import NaiveSyncBatchNorm1d
        torch.manual_seed(0)
        xs = [torch.randn(n, 16) * (1 + r) + r for r, n in enumerate((37, 91))]      # ragged per-rank batches
        bn = NaiveSyncBatchNorm1d(16, eps=1e-3, momentum=0.01).train()
        with torch.no_grad():
            bn.weight.copy_(torch.linspace(0.5, 1.5, 16))
            bn.bias.copy_(torch.linspace(-0.2, 0.2, 16))
        x = xs[rank].clone().requires_grad_(True)
        y = bn(x)
        # oracle: the reference's equal-weight rank average of [mean, meansqr] (mmdet3d/ops/norm.py:66-73)
        other = xs[1 - rank]
        stats_other = torch.cat([other.mean(0), (other * other).mean(0)])
        ref = O.batch_norm_train(xs[rank], bn.weight.detach(), bn.bias.detach(), 1e-3, sync=[stats_other])
        ok_fwd = torch.allclose(y, ref, rtol=1e-5, atol=1e-6)
        # backward: gradient of sum over BOTH ranks' outputs w.r.t. this rank's input (all_reduce in backward)
        (y * torch.arange(16.0)).sum().backward()
        xa = [t.clone().requires_grad_(True) for t in xs]
        mean = sum(t.mean(0) for t in xa) / world
        msq = sum((t * t).mean(0) for t in xa) / world
        scale = bn.weight.detach() * torch.rsqrt(msq - mean * mean + 1e-3)
        total = sum(((t * scale + (bn.bias.detach() - mean * scale)) * torch.arange(16.0)).sum() for t in xa)
        total.backward()
        ok_bwd = torch.allclose(x.grad, xa[rank].grad, rtol=1e-4, atol=1e-6)
        mean_all = sum(t.mean(0) for t in xs) / world
        ok_run = torch.allclose(bn.running_mean, 0.01 * mean_all, rtol=1e-5, atol=1e-7)
        # flat-gradient convention: all_reduce(sum) then a 1/world factor folded into the optimiser
        g = torch.full((5,), float(rank + 1))
        dist.all_reduce(g)
        ok_grad = torch.allclose(g / world, torch.full((5,), 1.5))
        ret[rank] = (ok_fwd, ok_bwd, ok_run, ok_grad)
    finally:
        dist.destroy_process_group()


def test_sync_bn_and_grad_allreduce_two_ranks():
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
        for rank in (0, 1):
            assert ret[rank] == (True, True, True, True), (rank, ret[rank])


def test_frame_sharding_is_disjoint_and_seeded():
    from bench import make_batches
    a = make_batches(0, 1, 2, dict(sweeps=1))[0]
    b = make_batches(1, 1, 2, dict(sweeps=1))[0]
    a2 = make_batches(0, 1, 2, dict(sweeps=1))[0]
    assert all((x == y).all() for x, y in zip(a, a2))                 # deterministic per rank
    assert not any(x.shape == y.shape and (x == y).all() for x in a for y in b)   # ranks see different frames
