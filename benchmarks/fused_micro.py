#!/usr/bin/env python
"""Micro-benchmark of the fused FiLM + compress forward kernel against the two separate kernels (development aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import robot_gnn_b200 as R

L = R._lib
G, K, F, Fo = 4096, 5, 512, 512
g = R.robot_batch(G, K, device="cuda"); ix = g.index()
gen = torch.Generator(device="cuda").manual_seed(0)
h = torch.rand((G * K, F, 64), generator=gen, device="cuda").bfloat16()
z = torch.rand((ix.E, 2 * F), generator=gen, device="cuda").bfloat16()
w = (torch.randn((Fo, 2 * F), generator=gen, device="cuda") * 0.03).bfloat16()
bias = torch.randn(Fo, device="cuda")
y = torch.empty((G * K, Fo, 64), dtype=torch.bfloat16, device="cuda")
gh = torch.empty_like(h)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def fused(store):
    L.check(L.lib().rgnn_tc_conv_film_fused(L.ptr(w), L.ptr(h), L.ptr(z), L.ptr(bias), L.ptr(y), L.ptr(gh) if store else None, G, K, F, Fo, L.stream()), "fused")


def unfused():
    R.functional._aggregate(ix, h, z, True, None, L.POST_MEAN, None, None, gh, None, False)
    L.check(L.lib().rgnn_tc_conv1x1(L.ptr(w), L.ptr(h), F, L.ptr(gh), F, L.ptr(y), Fo, None, 0, L.ptr(bias), G * K, L.stream()), "conv")


print("debug", os.environ.get("RGNN_FUSED_DEBUG", "0"), "fused+store %.3f ms | fused eval %.3f ms | unfused %.3f ms" %
      (timeit(lambda: fused(True)), timeit(lambda: fused(False)), timeit(unfused)))
