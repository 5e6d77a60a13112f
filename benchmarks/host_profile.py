"""Where the HOST time of a launch-bound step goes (script shape: 8 graphs x 5 robots, F = 512, bf16, two layers + compress):
cProfile of 300 eager training steps through gnn_slice, sorted by own time.  Usage: python benchmarks/host_profile.py [top_n]"""
import cProfile
import os
import pstats
import sys
import time
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import robot_gnn_b200 as R
    dev = torch.device("cuda")
    F, K, G = 512, 5, 8
    torch.manual_seed(1)
    blk = R.GNNBlock(types.SimpleNamespace(feature_dim=F, compress_gcn=True, multi_gcn=True)).to(dev)
    g = R.robot_batch(G, K, device=dev)
    g.index()
    N, E = G * K, G * K * (K - 1)
    h = torch.relu(torch.randn((N, F, 8, 8), device=dev)).bfloat16()
    dy = torch.randn((N, F, 8, 8), device=dev).bfloat16()
    g.edata["pose"] = torch.randn((E, 9), device=dev)

    params = list(blk.parameters())

    def step():
        hh = h.detach().requires_grad_()
        for p in params:
            p.grad = None
        R.gnn_slice(blk, g, hh).backward(dy)
    for _ in range(20):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(300):
        step()
    torch.cuda.synchronize()
    print(f"eager: {(time.perf_counter() - t0) / 300 * 1e3:.4f} ms/step")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(300):
        step()
    torch.cuda.synchronize()
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(int(sys.argv[1]) if len(sys.argv) > 1 else 30)


if __name__ == "__main__":
    main()
