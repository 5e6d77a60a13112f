"""Times the two CUDA-core kernels of the edge encoder's first layer (rgnn_edge_mlp1_fwd / _bwd) at the bench shape
(E = 81920 edges, 128 hidden features, bf16) and checks them against torch."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import robot_gnn_b200 as R
    L = R._lib
    E, H = int(os.environ.get("E", 81920)), int(os.environ.get("H", 128))
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(3)
    pose = torch.randn((E, 9), generator=g, device=dev)
    w1, b1 = torch.randn((H, 9), generator=g, device=dev) * 0.3, torch.randn((H,), generator=g, device=dev) * 0.1
    a1 = torch.empty((E, H), dtype=torch.bfloat16, device=dev)
    da1 = torch.randn((E, H), generator=g, device=dev).bfloat16()
    dw1, db1 = torch.empty((H, 9), device=dev), torch.empty((H,), device=dev)
    fwd = lambda: L.check(L.lib().rgnn_edge_mlp1_fwd(L.BF16, L.ptr(pose), L.ptr(w1), L.ptr(b1), L.ptr(a1), E, H, L.stream()), "fwd")
    bwd = lambda: L.check(L.lib().rgnn_edge_mlp1_bwd(L.BF16, L.ptr(pose), L.ptr(da1), L.ptr(dw1), L.ptr(db1), E, H, L.stream()), "bwd")
    for name, fn in (("mlp1_fwd", fwd), ("mlp1_bwd", bwd)):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(50):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"{name}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per launch")
    ref = torch.relu(pose @ w1.t() + b1)
    print("fwd max err", float((a1.float() - ref).abs().max()), "dw1 rel err",
          float((dw1 - da1.float().t() @ pose).abs().max() / (da1.float().t() @ pose).abs().max()),
          "db1 rel err", float((db1 - da1.float().sum(0)).abs().max() / da1.float().sum(0).abs().max()))


if __name__ == "__main__":
    main()
