"""Where does the host time of one step go at the reference's real batch (G=8 graphs of K=5 robots, dgl/training.py:28)?
cProfile over eager fwd+bwd steps (launch-bound: the GPU work is ~0.2 ms), plus the step time and launch count.
    python benchmarks/latency_profile.py [--feat 512] [--graphs 8] [--steps 300] [--no-profile]"""
import argparse
import cProfile
import io
import os
import pstats
import sys
import time
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import robot_gnn_b200 as R  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--feat", type=int, default=512)
    ap.add_argument("--graphs", type=int, default=8)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--eval", action="store_true", help="forward only under torch.no_grad()")
    a = ap.parse_args()
    dev = torch.device("cuda")
    G, K, F = a.graphs, 5, a.feat
    N, E = G * K, G * K * (K - 1)
    torch.manual_seed(1)
    blk = R.GNNBlock(types.SimpleNamespace(feature_dim=F, compress_gcn=True, multi_gcn=True)).to(dev)
    h = torch.relu(torch.randn((N, F, 8, 8), device=dev)).bfloat16().requires_grad_()
    dy = torch.randn((N, F, 8, 8), device=dev).bfloat16()
    pose = torch.randn((E, 9), device=dev)
    g = R.robot_batch(G, K, device=dev)
    g.index()

    def step():
        if a.eval:
            with torch.no_grad():
                return blk(g, h, pose)
        h.grad = None
        for p in blk.parameters():
            p.grad = None
        y = blk(g, h, pose)
        y.backward(dy)
        return y

    for _ in range(20):
        step()
    torch.cuda.synchronize()
    n0 = R._lib.lib().rgnn_launch_count()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / a.steps
    launches = (R._lib.lib().rgnn_launch_count() - n0) / a.steps
    print(f"G={G} F={F} {'eval' if a.eval else 'train'}: {dt * 1e3:.3f} ms/step, {launches:.1f} launches/step, {G / dt:.0f} graphs/s")
    if a.no_profile:
        return
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(a.steps):
        step()
    torch.cuda.synchronize()
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
    print("\n".join(l[:150] for l in s.getvalue().splitlines()[:45]))


if __name__ == "__main__":
    main()
