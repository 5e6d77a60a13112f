#!/usr/bin/env python
"""BASELINE.json configs[4]: scaling sweep over robots-per-graph x graphs-per-batch (fwd+bwd of the two-layer block, F=512,
bf16, one GPU per process; run under torchrun for the multi-GPU variant).  The 64 robots x 64k graphs corner is 0.67 TB of
bf16 features - cells whose tensors would exceed --mem-gb are skipped and reported as capped.

    python benchmarks/sweep.py [--robots 4 8 16 32 64] [--graphs 1024 4096 16384 65536] [--feat 512] [--mem-gb 120]
Prints one JSON line per cell.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def cell_bytes(g, k, f, es=2):
    n, e = g * k, g * k * (k - 1)
    feat = n * f * 64 * es
    edge = e * f * es
    # per layer: h, agg, y, dy, dh, dgh (features) ; a1, da1 (edge x F) ; z, dz, dzp (edge x 2F)
    return 2 * (6 * feat + 2 * edge + 3 * 2 * edge) + 2 * feat


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--robots", type=int, nargs="+", default=[4, 8, 16, 32, 64])
    ap.add_argument("--graphs", type=int, nargs="+", default=[1024, 4096, 16384, 65536])
    ap.add_argument("--feat", type=int, default=512)
    ap.add_argument("--mem-gb", type=float, default=120.0)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    import types
    import torch.distributed as dist
    import robot_gnn_b200 as R
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    world, rank = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0))
    if world > 1:       # --graphs is the per-GPU batch (weak scaling); one NCCL all-reduce of the GNN-layer gradients per step
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1)
    blk = R.GNNBlock(types.SimpleNamespace(feature_dim=a.feat, compress_gcn=True, multi_gcn=True))
    blk.to(dev)
    sync = R.parallel.GradSync(blk) if world > 1 else None
    for k in a.robots:
        for g in a.graphs:
            need = cell_bytes(g, k, a.feat) / 1e9
            rec = {"robots": k, "graphs": g, "feat": a.feat, "est_gb": round(need, 1)}
            if need > a.mem_gb:
                rec["capped"] = True
                if rank == 0:
                    print(json.dumps(rec), flush=True)
                continue
            n, e = g * k, g * k * (k - 1)
            gen = torch.Generator(device=dev).manual_seed(1 + rank)
            h = torch.rand((n, a.feat, 8, 8), generator=gen, device=dev).bfloat16().requires_grad_()
            dy = torch.randn((n, a.feat, 8, 8), generator=gen, device=dev).bfloat16()
            pose = torch.randn((e, 9), generator=gen, device=dev)
            gr = R.robot_batch(g, k, device=dev)

            def step():
                h.grad = None
                if sync is not None:
                    sync.zero_grad()
                else:
                    for p in blk.parameters():
                        p.grad = None
                blk(gr, h, pose).backward(dy)
                if sync is not None:
                    sync.finish()

            def barrier():
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
            for _ in range(3):
                step()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.steps):
                step()
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            flops = 2 * 3 * (4 * e * a.feat ** 2 + 4 * n * 64 * a.feat ** 2)
            rec.update(n_gpus=world, graphs_per_gpu=g, ms_per_step=round(ms, 3), graphs_per_s=round(world * g / ms * 1e3, 1),
                       tflops_per_gpu=round(flops / ms / 1e9, 1),
                       path="streaming ring kernels + fused dX/FiLM backward (K<=6..8)" if k <= 8 else
                            ("mma.sync per-channel K x K products (K % 16 == 0)" if k % 16 == 0 else "dense CUDA-core complete-graph kernels"))
            if sync is not None:
                rec["grad_sync_check"] = "ok" if sync.check_ranks_agree()[0] else "MISMATCH"
            if rank == 0:
                print(json.dumps(rec), flush=True)
            del h, dy, pose, gr
            torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
