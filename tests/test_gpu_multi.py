"""GPU, 2 ranks over NCCL (skipped with fewer than 2 devices): SURVEY §4 'multi-GPU: 1-vs-N-rank gradient equality after
all-reduce on an identical global batch'.  Each rank owns a contiguous half of the graphs, builds its own CSR, exchanges
nothing but the bucketed all-reduce of the GNN-layer gradients."""
import os
import socket
import sys
import types

import numpy as np
import pytest
import torch

from conftest import ROOT, nerr

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import robot_gnn_b200 as R
    from robot_check import make_block
    from oracle import gnn_oracle as O
    G, K, F = 16, 5, 128
    blk, _ = make_block(R, O, F, True, True)
    blk.cuda()
    sync = R.parallel.GradSync(blk, average=False)          # sum == gradient of the summed loss over the global batch
    sb = O.synthetic_batch(G, K, F, p=64, seed=5)
    lo, hi = R.parallel.shard_range(G, rank, world)
    n0, n1, e0, e1 = lo * K, hi * K, lo * K * (K - 1), hi * K * (K - 1)
    g = R.robot_batch(hi - lo, K, device="cuda")
    h = sb["h"].reshape(G * K, F, 8, 8)[n0:n1].cuda().requires_grad_()
    y = blk(g, h, sb["pose"][e0:e1].cuda())
    y.backward(sb["dy"].reshape(G * K, F, 8, 8)[n0:n1].cuda())
    sync.finish()
    torch.cuda.synchronize()
    q.put((rank, {n: p.grad.cpu().numpy() for n, p in blk.named_parameters()}, y.detach().cpu().numpy(), h.grad.cpu().numpy()))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_rank_gradients_equal_single_rank(R):
    import torch.multiprocessing as mp
    from robot_check import make_block
    from oracle import gnn_oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    G, K, F = 16, 5, 128
    blk, _ = make_block(R, O, F, True, True)
    blk.cuda()
    sb = O.synthetic_batch(G, K, F, p=64, seed=5)
    g = R.robot_batch(G, K, device="cuda")
    h = sb["h"].reshape(G * K, F, 8, 8).cuda().requires_grad_()
    y = blk(g, h, sb["pose"].cuda())
    y.backward(sb["dy"].reshape(G * K, F, 8, 8).cuda())
    (_, g0, y0, dh0), (_, g1, y1, dh1) = res
    assert nerr(torch.from_numpy(np.concatenate([y0, y1])), y) <= 1e-5
    assert nerr(torch.from_numpy(np.concatenate([dh0, dh1])), h.grad) <= 1e-5
    for n, p in blk.named_parameters():
        assert np.array_equal(g0[n], g1[n]), n                        # every rank holds the same reduced gradient
        assert nerr(torch.from_numpy(g0[n]), p.grad) <= 2e-5, n       # == single-rank gradient of the whole batch
