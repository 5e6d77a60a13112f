#!/usr/bin/env python
"""Micro-benchmark of the FiLM aggregation kernels (forward and fused backward) at the bench shape: alone, and interleaved with a
tensor-core GEMM the way the training step runs them (the GEMM drives the GPU into its power cap; the memory-bound kernel that
follows then runs at lower clocks, which is what bench.py's per-kernel numbers see).  Development aid."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import robot_gnn_b200 as R

L = R._lib
G, K, F = int(os.environ.get("G", 4096)), int(os.environ.get("K", 5)), int(os.environ.get("F", 512))
g = R.robot_batch(G, K, device="cuda"); ix = g.index()
gen = torch.Generator(device="cuda").manual_seed(0)
h = torch.rand((G * K, F, 64), generator=gen, device="cuda").bfloat16()
z = torch.rand((ix.E, 2 * F), generator=gen, device="cuda").bfloat16()
dcat = torch.randn((G * K, 2 * F, 64), generator=gen, device="cuda").bfloat16()
w = (torch.randn((F, 2 * F), generator=gen, device="cuda") * 0.03).bfloat16()
bias = torch.randn(F, device="cuda")
y = torch.empty((G * K, F, 64), dtype=torch.bfloat16, device="cuda")
gh = torch.empty_like(h)
dh = torch.empty_like(h)
HBM = 6547.5


def fwd():
    R.functional._aggregate(ix, h, z, True, None, L.POST_MEAN, None, None, gh, None, False)


def bwd():
    R.functional._film_backward(ix, h, dcat[:, F:], z, dcat[:, :F], True, ldd=2 * F * 64, ldadd=2 * F * 64, dh=dh)


def gemm():
    L.check(L.lib().rgnn_tc_conv1x1(L.ptr(w), L.ptr(h), F, L.ptr(gh), F, L.ptr(y), F, None, 0, L.ptr(bias), G * K, L.stream()), "conv")


def timeit(fn, n=20, before=None):
    for _ in range(3):
        if before: before()
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(n):
        if before: before()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in evs)
    return t[len(t) // 2]


es = 2
fb = es * (2 * G * K * F * 64 + 2 * ix.E * F)
bb = es * (4 * G * K * F * 64 + 4 * ix.E * F)
for name, fn, nb in (("film fwd", fwd, fb), ("film bwd", bwd, bb)):
    a = timeit(fn); b = timeit(fn, before=gemm)
    print(f"{name}: alone {a:.3f} ms ({nb / a / 1e6:.0f} GB/s, {nb / a / 1e6 / HBM:.2f}) | after a GEMM {b:.3f} ms ({nb / b / 1e6:.0f} GB/s, {nb / b / 1e6 / HBM:.2f})")
