#!/usr/bin/env python
"""North-star extra (parity unpinned by the reference): a stack of DGL-formula GCN2Conv layers (GCNII: initial residual +
identity mapping, ((1-a) A_hat X + a X0)((1-b) I + b W)) over batched robot graphs, channels-last feature maps [N, P, F] in bf16,
fwd+bwd.  The sym-normalised aggregation runs in the robot-graph aggregation kernels, the weight transform on the tcgen05 GEMM with
the identity mapping fused in its epilogue.  Prints one JSON line.

    python benchmarks/gcnii.py [--graphs 4096] [--robots 5] [--feat 512] [--layers 2]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--graphs", type=int, default=4096)
    ap.add_argument("--robots", type=int, default=5)
    ap.add_argument("--feat", type=int, default=512)
    ap.add_argument("--pixels", type=int, default=64)
    ap.add_argument("--layers", type=int, default=2)
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    import robot_gnn_b200 as R
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    torch.manual_seed(1)
    layers = [R.GCN2Conv(a.feat, layer=i + 1, alpha=0.1, lambda_=1).to(dev) for i in range(a.layers)]
    for l in layers:
        torch.nn.init.normal_(l.weight1, std=a.feat ** -0.5)
    g = R.robot_batch(a.graphs, a.robots, device=dev)
    n = a.graphs * a.robots
    gen = torch.Generator(device=dev).manual_seed(1)
    x0 = torch.randn((n, a.pixels, a.feat), generator=gen, device=dev).bfloat16().requires_grad_()
    dy = torch.randn((n, a.pixels, a.feat), generator=gen, device=dev).bfloat16()

    def step():
        x0.grad = None
        for l in layers:
            for p in l.parameters():
                p.grad = None
        x = x0
        for l in layers:
            x = l(g, x, x0)
        x.backward(dy)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    R._lib.PROFILE = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    prof = R._lib.profile_summary()
    R._lib.PROFILE = None
    ms = e0.elapsed_time(e1) / a.steps
    rows = n * a.pixels
    flops = a.layers * 3 * 2.0 * rows * a.feat * a.feat
    print(json.dumps({"workload": f"GCNII stack: {a.layers} x GCN2Conv, [N={n}, P={a.pixels}, F={a.feat}] bf16, fwd+bwd",
                      "ms_per_step": round(ms, 3), "graphs_per_s": round(a.graphs / ms * 1e3, 1), "gemm_tflops": round(flops / ms / 1e9, 1),
                      "kernels": {k: {"calls": v["calls"], "ms_avg": round(v["ms_avg"], 4),
                                      "achieved": round(v["work_total"] / (v["ms_total"] * 1e-3) / (1e9 if k.startswith(("aggregate", "mean_mix")) else 1e12), 1),
                                      "unit": "GB/s" if k.startswith(("aggregate", "mean_mix")) else "TFLOP/s"} for k, v in prof.items()}}))


if __name__ == "__main__":
    main()
