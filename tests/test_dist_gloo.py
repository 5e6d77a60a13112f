"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path - graph sharding and the
bucketed gradient all-reduce.  (Kernels are not involved; the N>1 GPU run is bench.py under torchrun.)"""
import os
import socket
import sys
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import robot_gnn_b200 as R
    torch.manual_seed(0)                                   # identical parameters on every rank
    blk = R.GNNBlock(types.SimpleNamespace(feature_dim=8, compress_gcn=True, multi_gcn=True))
    sync = R.parallel.GradSync(blk)
    assert len(sync.buckets) == 2 and sync.buckets[0][0][0].startswith(("gcn2", "conv2"))   # layer 2 reduces first
    # rank-dependent synthetic gradients, produced through autograd so the hooks fire in backward order
    # (the collective runs IN PLACE on the gradient buckets as soon as a bucket is complete, so the rank-local gradients are taken
    #  from a pass with the collective switched off)
    def local_loss():
        gen = torch.Generator().manual_seed(100 + rank)
        return sum((p * torch.randn(p.shape, generator=gen)).sum() for p in blk.parameters())
    sync.enabled = False
    local_loss().backward()
    sync.finish()
    local = {n: p.grad.clone().numpy() for n, p in blk.named_parameters()}
    sync.enabled = True
    sync.zero_grad()
    local_loss().backward()
    sync.finish()
    out = {n: p.grad.clone().numpy() for n, p in blk.named_parameters()}
    # second step must work too (state reset)
    for p in blk.parameters():
        p.grad = None
    loss = sum((p * 2.0).sum() for p in blk.parameters())
    loss.backward()
    sync.finish()
    ok2 = all(torch.allclose(p.grad, torch.full_like(p, 2.0)) for p in blk.parameters())
    # third step the intended way: zero_grad() keeps every .grad a view into its persistent flat bucket (no cat / copy-back),
    # autograd accumulates in place, and after finish() the ranks hold bit-identical buckets
    sync.zero_grad()
    views_ok = all(p.grad.data_ptr() == v.data_ptr() for b, vs in zip(sync.buckets, sync._views) for (_, p), v in zip(b, vs))
    loss = sum((p * float(rank + 1)).sum() for p in blk.parameters())
    loss.backward()
    views_ok = views_ok and all(p.grad.data_ptr() == v.data_ptr() for b, vs in zip(sync.buckets, sync._views) for (_, p), v in zip(b, vs))
    sync.finish()
    ok3 = all(torch.allclose(p.grad, torch.full_like(p, 1.5)) for p in blk.parameters())
    agree, _ = sync.check_ranks_agree()
    # accumulate over two backward passes with the hooks quiet, reduce once (HostPipeline chunks)
    sync.zero_grad()
    sync.overlap = False
    for _ in range(2):
        sum((p * float(rank + 1)).sum() for p in blk.parameters()).backward()
    sync.finish()
    ok4 = all(torch.allclose(p.grad, torch.full_like(p, 3.0)) for p in blk.parameters())
    sync.remove()
    # bf16 wire format: half the bytes on the link, fp32 master buckets
    for p in blk.parameters():
        p.grad = None
    sync16 = R.parallel.GradSync(blk, wire_dtype=torch.bfloat16)
    sum((p * (0.1 * (rank + 1))).sum() for p in blk.parameters()).backward()
    sync16.finish()
    ok5 = all(torch.allclose(p.grad, torch.full_like(p, 0.15), rtol=1e-2) for p in blk.parameters()) and sync16.check_ranks_agree()[0] \
        and sync16.wire_bytes() * 2 == sync16.grad_bytes()
    ok2 = ok2 and views_ok and ok3 and agree and ok4 and ok5
    lo, hi = R.parallel.shard_range(11, rank, world)
    q.put((rank, local, out, ok2, (lo, hi)))
    dist.destroy_process_group()


def test_grad_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, l0, o0, ok0, s0), (_, l1, o1, ok1, s1) = res
    assert ok0 and ok1
    for n in l0:
        assert np.allclose(o0[n], (l0[n] + l1[n]) / 2, atol=1e-6) and np.array_equal(o0[n], o1[n]), n
    assert s0 == (0, 6) and s1 == (6, 11)


def test_shard_range_partitions():
    import robot_gnn_b200 as R
    for g, w in ((8, 8), (1000, 8), (7, 4), (3, 8)):
        cuts = [R.parallel.shard_range(g, r, w) for r in range(w)]
        assert cuts[0][0] == 0 and cuts[-1][1] == g
        assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))
        assert max(h - l for l, h in cuts) - min(h - l for l, h in cuts) <= 1
