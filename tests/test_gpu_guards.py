"""GPU: out-of-bounds self-check (compute-sanitizer is closed on this pool).  Every output of the hot kernels is a view into a
larger buffer with sentinel-filled guard bands on both sides; after the call the guards must be untouched and the payload fully
written (no sentinel left).  Shapes are chosen to hit tails: robots not a multiple of the tile's 4, rows not a multiple of 128,
ragged K, K > 8."""
import pytest
import torch

pytestmark = pytest.mark.gpu

GUARD = 4096     # elements on each side


class Guarded:
    def __init__(self, shape, dtype, sentinel):
        n = 1
        for s in shape:
            n *= s
        self.n, self.sentinel = n, sentinel
        self.buf = torch.full((n + 2 * GUARD,), sentinel, dtype=dtype, device="cuda")
        self.view = self.buf[GUARD:GUARD + n].view(shape)

    def check(self, name, allow_sentinel_inside=False):
        torch.cuda.synchronize()
        s = torch.tensor(self.sentinel, dtype=self.buf.dtype, device="cuda")
        assert bool((self.buf[:GUARD] == s).all()), f"{name}: write before the buffer"
        assert bool((self.buf[GUARD + self.n:] == s).all()), f"{name}: write past the buffer"
        if not allow_sentinel_inside:
            assert not bool((self.view == s).any()), f"{name}: part of the output was never written"


def bf(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(torch.bfloat16)


@pytest.mark.parametrize("nodes,c0,c1,m0,m1", [(5, 128, 128, 128, 0), (9, 64, 0, 128, 128), (3, 256, 256, 384, 0)])
def test_conv1x1_guards(R, nodes, c0, c1, m0, m1):
    L = R._lib
    w = bf(m0 + m1, c0 + c1, seed=1, scale=0.05)
    x0, x1 = bf(nodes, c0, 64, seed=2), (bf(nodes, c1, 64, seed=3) if c1 else None)
    o0 = Guarded((nodes, m0, 64), torch.bfloat16, 777.0)
    o1 = Guarded((nodes, m1, 64), torch.bfloat16, 777.0) if m1 else None
    L.check(L.lib().rgnn_tc_conv1x1(L.ptr(w), L.ptr(x0), c0, L.ptr(x1), c1, L.ptr(o0.view), m0, L.ptr(o1.view) if o1 else None, m1,
                                    None, nodes, L.stream()), "conv")
    o0.check("conv out0")
    if o1:
        o1.check("conv out1")
    # values too (odd M-tile counts leave the second CTA of the last pair without a tile)
    x = torch.cat([x0] + ([x1] if c1 else []), 1).float()
    ref = torch.einsum("ok,vkp->vop", w.float(), x)
    got = torch.cat([o0.view] + ([o1.view] if o1 else []), 1).float()
    assert float((got - ref).abs().max()) <= 2e-2 * float(ref.abs().max())


@pytest.mark.parametrize("m,n,k", [(77, 128, 64), (300, 256, 128), (129, 384, 64)])
def test_gemm_tn_and_nn_guards(R, m, n, k):
    L = R._lib
    a, b = bf(m, k, seed=4), bf(n, k, seed=5, scale=0.1)
    c = Guarded((m, n), torch.bfloat16, 555.0)
    L.check(L.lib().rgnn_tc_gemm_tn(L.ptr(a), k, L.ptr(b), k, L.ptr(c.view), n, m, n, k, 0, None, None, 1.0, 0.0, L.stream()), "tn")
    c.check("gemm_tn")
    ref = a.float() @ b.float().t()
    assert float((c.view.float() - ref).abs().max()) <= 2e-2 * float(ref.abs().max())
    kk = 130
    at, bt = bf(kk, 256, seed=6), bf(kk, n, seed=7)
    cf = Guarded((256, n), torch.float32, 9e9)
    part = Guarded((3, 256, n), torch.float32, 9e9)
    L.check(L.lib().rgnn_tc_gemm_nn(L.ptr(at), 256, L.ptr(bt), n, L.ptr(cf.view), n, 256, n, kk, 1.0, L.ptr(part.view), 3, L.stream()), "nn")
    cf.check("gemm_nn C")
    part.check("gemm_nn partial")


def test_wgrad_guards(R):
    L = R._lib
    nodes, m, c0, c1, splits = 11, 128, 256, 256, 4
    dy, x0, x1 = bf(nodes, m, 64, seed=8), bf(nodes, c0, 64, seed=9), bf(nodes, c1, 64, seed=10)
    dw = Guarded((m, c0 + c1), torch.float32, 9e9)
    db = Guarded((m,), torch.float32, 9e9)
    part = Guarded((splits * m * (c0 + c1 + 1),), torch.float32, 9e9)
    L.check(L.lib().rgnn_tc_conv1x1_wgrad(L.ptr(dy), m, L.ptr(x0), c0, L.ptr(x1), c1, L.ptr(dw.view), c0 + c1, L.ptr(db.view),
                                          L.ptr(part.view), splits, nodes, L.stream()), "wgrad")
    dw.check("wgrad dW")
    db.check("wgrad dbias")
    part.check("wgrad partial", allow_sentinel_inside=True)      # splits may be trimmed: unused partial rows stay untouched


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("F", [24, 64])        # 64: the streaming ring kernels (whole 16 / 32-channel items); 24: the register-prefetch ones
@pytest.mark.parametrize("ks", [[5] * 7, [12] * 3, [3, 5, 2, 4, 1], [20] * 2, [8] * 3, [2] * 5])
def test_aggregation_guards(R, dtype, ks, F):
    import numpy as np
    from oracle import gnn_oracle as O
    ed = [O.fully_connected_edges(k) for k in ks]
    g = R.RobotGraph(torch.from_numpy(np.concatenate([e[0] for e in ed])).cuda(), torch.from_numpy(np.concatenate([e[1] for e in ed])).cuda(),
                     ks, [len(e[0]) for e in ed], complete_k=ks[0] if len(set(ks)) == 1 else 0)
    ix = g.index()
    n, e = ix.N, ix.E
    gen = torch.Generator(device="cuda").manual_seed(0)
    h = torch.rand((n, F, 64), generator=gen, device="cuda").to(dtype)
    z = torch.rand((e, 2 * F), generator=gen, device="cuda").to(dtype)
    d = torch.randn((n, F, 64), generator=gen, device="cuda").to(dtype)
    L = R._lib
    Fn = R.functional
    out = Guarded((n, F, 64), dtype, 333.0)
    Fn._aggregate(ix, h, z, True, None, L.POST_MEAN, None, None, out.view, None, False)
    out.check("aggregate fwd")
    cat = Guarded((n, 2 * F, 64), dtype, 333.0)
    Fn._aggregate(ix, h, z, True, None, L.POST_MEAN, None, None, cat.view.view(-1)[F * 64:], cat.view, False, ldout=2 * F * 64, shape=(F, 64))
    cat.check("aggregate fwd (cat)")
    dh = Guarded((n, F, 64), dtype, 333.0)
    dz = Guarded((e, 2 * F), dtype, 333.0)
    rc = L.lib().rgnn_film_backward(L.dtype_code(h), L.ptr(h), F * 64, L.ptr(d), F * 64, L.ptr(z), 1, None, 0, L.ptr(dh.view), F * 64,
                                    L.ptr(dz.view), L.ptr(ix.inv_in_deg), L.ptr(ix.ptr32), L.ptr(ix.idx32), L.ptr(ix.noff32), ix.G,
                                    ix.max_nodes, ix.max_edges, ix.complete_k, n, e, F, 64, L.stream())
    assert rc == 0
    dh.check("film backward dh")
    dz.check("film backward dz", allow_sentinel_inside=(e == 0))
    # the layout the fused layer uses: dAgg and the direct gradient are the two channel halves of one [N, 2F, 64] buffer
    dcat = torch.randn((n, 2 * F, 64), generator=gen, device="cuda").to(dtype)
    dh2 = Guarded((n, F, 64), dtype, 333.0)
    dz2 = Guarded((e, 2 * F), dtype, 333.0)
    rc = L.lib().rgnn_film_backward(L.dtype_code(h), L.ptr(h), F * 64, L.ptr(dcat.view(-1)[F * 64:]), 2 * F * 64, L.ptr(z), 1, L.ptr(dcat), 2 * F * 64,
                                    L.ptr(dh2.view), F * 64, L.ptr(dz2.view), L.ptr(ix.inv_in_deg), L.ptr(ix.ptr32), L.ptr(ix.idx32),
                                    L.ptr(ix.noff32), ix.G, ix.max_nodes, ix.max_edges, ix.complete_k, n, e, F, 64, L.stream())
    assert rc == 0
    dh2.check("film backward dh (strided halves)")
    dz2.check("film backward dz (strided halves)", allow_sentinel_inside=(e == 0))
    # and it must equal the un-fused composition: addend + backward without addend on a contiguous copy of the dAgg half
    d2 = dcat[:, F:].contiguous()
    dh3, dz3 = torch.empty_like(h), torch.empty_like(z)
    rc = L.lib().rgnn_film_backward(L.dtype_code(h), L.ptr(h), F * 64, L.ptr(d2), F * 64, L.ptr(z), 1, None, 0, L.ptr(dh3), F * 64,
                                    L.ptr(dz3), L.ptr(ix.inv_in_deg), L.ptr(ix.ptr32), L.ptr(ix.idx32), L.ptr(ix.noff32), ix.G,
                                    ix.max_nodes, ix.max_edges, ix.complete_k, n, e, F, 64, L.stream())
    assert rc == 0
    assert torch.equal(dz3, dz2.view)
    ref = dh3.float() + dcat[:, :F].float()
    assert float((dh2.view.float() - ref).abs().max()) <= (1e-5 if dtype == torch.float32 else 6e-2) * float(ref.abs().max())


def test_csr_guards(R):
    import numpy as np
    from oracle import gnn_oracle as O
    ks = [4, 1, 7, 2]
    ed = [O.fully_connected_edges(k) for k in ks]
    g = R.RobotGraph(torch.from_numpy(np.concatenate([e[0] for e in ed])).cuda(), torch.from_numpy(np.concatenate([e[1] for e in ed])).cuda(),
                     ks, [len(e[0]) for e in ed])
    src, dst, noff, eoff = g._globals()
    n, e, G = g.num_nodes(), g.num_edges(), len(ks)
    L = R._lib
    i64 = lambda k: Guarded((k,), torch.int64, -12345)
    arrs = dict(in_deg=i64(n), out_deg=i64(n), in_ptr=i64(n + 1), in_src=i64(e), in_eid=i64(e), out_ptr=i64(n + 1), out_dst=i64(e), out_eid=i64(e))
    ptr32 = Guarded((2 * (n + 1) + G + 1,), torch.int32, -777)
    idx32 = Guarded((6 * e,), torch.int32, -777)
    norm = Guarded((3 * n,), torch.float32, 9e9)
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    L.check(L.lib().rgnn_build_batched_csr(L.ptr(src), L.ptr(dst), n, e, L.ptr(noff), L.ptr(eoff), G,
                                           *[L.ptr(arrs[k].view) for k in ("in_deg", "out_deg", "in_ptr", "in_src", "in_eid", "out_ptr", "out_dst", "out_eid")],
                                           L.ptr(ptr32.view), L.ptr(idx32.view), L.ptr(norm.view), L.ptr(status), L.stream()), "csr")
    for k, a in arrs.items():
        a.check(k)
    ptr32.check("ptr32"); idx32.check("idx32"); norm.check("norm")
    assert int(status.item()) == 0
