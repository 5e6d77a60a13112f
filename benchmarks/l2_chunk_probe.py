"""Can the 126 MB L2 carry the intermediates of one layer between its kernels?  (north-star (c): aggregated features / dcat should
not round-trip HBM.)  The batch is cut into chunks of whole graphs and each chunk runs  producer kernel -> consumer kernel  back
to back, so the consumer finds the producer's output (and the shared input h) in L2:
    forward :  FiLM-mean(h, z) -> g_h      then  conv1x1([h ; g_h]) -> y
    backward:  conv dX(dy) -> dh_direct, dg_h   then  FiLM backward(h, dg_h, z, dh_direct) -> dh, dz
The chunk loop is captured into a CUDA graph (at 32+ chunks the Python launch path would be the bottleneck) and timed against
the unchunked pair.  Prints one JSON line per chunk size.
    python benchmarks/l2_chunk_probe.py [--graphs 4096] [--feat 512]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import robot_gnn_b200 as R  # noqa: E402
from robot_gnn_b200 import functional as Fn  # noqa: E402

L = R._lib


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--graphs", type=int, default=4096)
    ap.add_argument("--feat", type=int, default=512)
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    dev = torch.device("cuda")
    G, K, F, P = a.graphs, 5, a.feat, 64
    N, E = G * K, G * K * (K - 1)
    gen = torch.Generator(device=dev).manual_seed(1)
    bf = lambda *s: torch.randn(s, generator=gen, device=dev).bfloat16()
    h, dy = torch.relu(bf(N, F, P)), bf(N, F, P)
    z = torch.sigmoid(bf(E, 2 * F).float()).bfloat16()
    wc = (torch.rand((F, 2 * F), generator=gen, device=dev) * 2 - 1) / (2 * F) ** 0.5
    wb, wt = wc.bfloat16().contiguous(), wc.t().contiguous().bfloat16()
    bias = torch.zeros(F, device=dev)
    gh, y = torch.empty_like(h), torch.empty_like(h)
    dhd, dgh, dh = torch.empty_like(h), torch.empty_like(h), torch.empty_like(h)
    dz = torch.empty_like(z)
    inv = torch.full((N,), 1.0 / (K - 1), device=dev)
    g = R.robot_batch(G, K, device=dev)
    ix = g.index()
    lib, ptr, st = L.lib(), L.ptr, L.stream
    es = 2

    def off(t, rows, row_elems):
        import ctypes
        return ctypes.c_void_p(t.data_ptr() + rows * row_elems * es)

    def fwd(g0, gc):
        n0, nc, e0 = g0 * K, gc * K, g0 * K * (K - 1)
        L.check(lib.rgnn_aggregate(L.BF16, off(h, n0, F * P), F * P, off(z, e0, 2 * F), 1, None, L.POST_MEAN, None, None, F * P, off(gh, n0, F * P), F * P, None,
                                   ptr(ix.in_ptr32), ptr(ix.in_src32), ptr(ix.in_eid32), ptr(ix.noff32), gc, K, K * (K - 1), K, nc, gc * K * (K - 1), F, P, st()), "agg")
        L.check(lib.rgnn_tc_conv1x1(ptr(wb), off(h, n0, F * P), F, off(gh, n0, F * P), F, off(y, n0, F * P), F, None, 0, ptr(bias), nc, st()), "conv")

    def bwd(g0, gc):
        n0, nc, e0 = g0 * K, gc * K, g0 * K * (K - 1)
        L.check(lib.rgnn_tc_conv1x1(ptr(wt), off(dy, n0, F * P), F, None, 0, off(dhd, n0, F * P), F, off(dgh, n0, F * P), F, None, nc, st()), "dx")
        L.check(lib.rgnn_film_backward(L.BF16, off(h, n0, F * P), F * P, off(dgh, n0, F * P), F * P, off(z, e0, 2 * F), 1, off(dhd, n0, F * P), F * P,
                                       off(dh, n0, F * P), F * P, off(dz, e0, 2 * F), ptr(inv), ptr(ix.ptr32), ptr(ix.idx32), ptr(ix.noff32),
                                       gc, K, K * (K - 1), K, nc, gc * K * (K - 1), F, P, st()), "film_bwd")

    def timed_graph(fn, chunk):
        def run():
            for g0 in range(0, G, chunk):
                fn(g0, min(chunk, G - g0))
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            run()
        torch.cuda.current_stream().wait_stream(s)
        cg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(cg):
            run()
        for _ in range(3):
            cg.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            cg.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.reps

    ref_y = ref_dh = None
    for chunk in (G, 1024, 512, 256, 128, 64, 32):
        if chunk > G:
            continue
        tf, tb = timed_graph(fwd, chunk), timed_graph(bwd, chunk)
        if ref_y is None:
            ref_y, ref_dh, ref_dz = y.clone(), dh.clone(), dz.clone()
        same = bool(torch.equal(y, ref_y) and torch.equal(dh, ref_dh) and torch.equal(dz, ref_dz))
        mb = chunk * K * F * P * es / 1e6
        print(json.dumps(dict(chunk_graphs=chunk, chunks=(G + chunk - 1) // chunk, tensor_chunk_MB=round(mb, 1), fwd_ms=round(tf, 4), bwd_ms=round(tb, 4),
                              identical_to_unchunked=same)))


if __name__ == "__main__":
    main()
