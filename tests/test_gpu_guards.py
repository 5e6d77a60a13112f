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


# ------------------------------------------------------------------------------------------ round-2 kernels
@pytest.mark.parametrize("g_,k,f,fo", [(3, 5, 64, 64), (7, 5, 192, 128), (2, 2, 128, 64), (5, 6, 64, 192), (149, 3, 64, 64),
                                          (75, 5, 128, 128), (9, 4, 256, 64), (5, 6, 128, 192), (151, 3, 128, 64)])
def test_fused_dx_film_backward_guards(R, g_, k, f, fo):
    """rgnn_tc_conv_dx_film_bwd: dh / dz / the dz column sums are written completely and nothing else, incl. graph counts that are not a
    multiple of the SM count and widths with an odd number of channel tiles; values against the unfused formulas in fp32."""
    import subprocess, sys, os, textwrap
    # (a child process: a fault in the kernel under test must not take the test session's CUDA context with it)
    code = textwrap.dedent(f"""
        import sys, os, torch
        sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r}); sys.path.insert(0, {os.path.dirname(os.path.abspath(__file__))!r})
        import robot_gnn_b200 as R
        from test_gpu_guards import Guarded, bf
        L = R._lib
        g_, k, f, fo = {g_}, {k}, {f}, {fo}
        n, e = g_ * k, g_ * k * (k - 1)
        h, dy = bf(n, f, 64, seed=1), bf(n, fo, 64, seed=2)
        z = torch.sigmoid(bf(e, 2 * f, seed=3).float()).bfloat16()
        wt = bf(2 * f, fo, seed=4, scale=0.05)
        wtp = wt.index_select(0, R.functional._dx_film_perm(f, wt.device))
        dh, dz = Guarded((n, f, 64), torch.bfloat16, 777.0), Guarded((e, 2 * f), torch.bfloat16, 777.0)
        db2 = Guarded((2 * f,), torch.float32, 777.0)             # (overwritten, not accumulated into: the launcher zeroes it)
        L.check(L.lib().rgnn_tc_conv_dx_film_bwd(L.ptr(wtp), L.ptr(dy), L.ptr(h), L.ptr(z), 1, L.ptr(dh.view), L.ptr(dz.view), L.ptr(db2.view), g_, k, f, fo, L.stream()), "dxf")
        dh.check("dx_film_bwd dh"); dz.check("dx_film_bwd dz")
        # values against the unfused formulas in fp32
        dcat = torch.einsum("ko,vop->vkp", wt.float(), dy.float())
        dd, da = dcat[:, :f], dcat[:, f:]
        src = torch.arange(k).repeat_interleave(k); dst = torch.arange(k).repeat(k); keep = src != dst
        src, dst = src[keep].cuda(), dst[keep].cuda()
        gam = z.float()[:, 0::2].view(g_, k * (k - 1), f); bet = z.float()[:, 1::2].view(g_, k * (k - 1), f)
        dav, hv = da.view(g_, k, f, 64), h.float().view(g_, k, f, 64)
        w = 1.0 / (k - 1)
        msg = gam[..., None] * dav[:, dst] * w
        ref_dh = dd.view(g_, k, f, 64).clone().index_add_(1, src, msg)
        dgam = (hv[:, src] * dav[:, dst]).sum(-1) * w * gam * (1 - gam)
        dbet = dav[:, dst].sum(-1) * w * bet * (1 - bet)
        assert float((dh.view.float().view(g_, k, f, 64) - ref_dh).abs().max()) <= 2e-2 * float(ref_dh.abs().max())
        got = dz.view.float().view(g_, k * (k - 1), f, 2)
        assert float((got[..., 0] - dgam).abs().max()) <= 2e-2 * float(dgam.abs().max())
        assert float((got[..., 1] - dbet).abs().max()) <= 2e-2 * float(dbet.abs().max())
        ref_db2 = torch.stack((dgam.sum((0, 1)), dbet.sum((0, 1))), -1).reshape(-1)
        assert float((db2.view - ref_db2).abs().max()) <= 1e-3 * float(ref_db2.abs().max())
        db2.check("dx_film_bwd db2")
        print("ok")
    """)
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


@pytest.mark.parametrize("g_,k,f", [(3, 16, 8), (2, 32, 24), (1, 64, 16), (5, 48, 8)])
def test_film_mma_guards(R, g_, k, f):
    """K = 16..64 robots (mma.sync kernels): forward output, dh and dz written completely and nothing else."""
    L = R._lib
    n, e = g_ * k, g_ * k * (k - 1)
    g = R.robot_batch(g_, k, device="cuda")
    ix = g.index()
    h, dagg, add = bf(n, f, 64, seed=1), bf(n, f, 64, seed=2), bf(n, f, 64, seed=3)
    z = torch.sigmoid(bf(e, 2 * f, seed=4).float()).bfloat16()
    out = Guarded((n, f, 64), torch.bfloat16, 777.0)
    R.functional._aggregate(ix, h, z, True, None, L.POST_MEAN, None, None, out.view, None, False)
    out.check("film_mma fwd")
    dh, dz = Guarded((n, f, 64), torch.bfloat16, 777.0), Guarded((e, 2 * f), torch.bfloat16, 777.0)
    L.check(L.lib().rgnn_film_backward(L.BF16, L.ptr(h), f * 64, L.ptr(dagg), f * 64, L.ptr(z), 1, L.ptr(add), f * 64, L.ptr(dh.view), f * 64,
                                       L.ptr(dz.view), L.ptr(ix.inv_in_deg), L.ptr(ix.ptr32), L.ptr(ix.idx32), L.ptr(ix.noff32), ix.G, ix.max_nodes,
                                       ix.max_edges, ix.complete_k, ix.N, ix.E, f, 64, L.stream()), "film_backward")
    dh.check("film_mma dh")
    dz.check("film_mma dz")


@pytest.mark.parametrize("n,rows,cols", [(3, 49, 48), (1, 64, 130), (7, 9, 5)])
def test_layout_cast_and_split3_guards(R, n, rows, cols):
    L = R._lib
    x = torch.randn((n, rows, cols), device="cuda")
    for dt, code in ((torch.bfloat16, L.BF16), (torch.float32, L.F32)):
        o = Guarded((n, cols, rows), dt, 777.0)
        L.check(L.lib().rgnn_layout_cast(L.F32, L.ptr(x), code, L.ptr(o.view), n, rows, cols, L.stream()), "layout_cast")
        o.check("layout_cast")
        assert torch.equal(o.view, x.transpose(1, 2).to(dt))
    m = 4 * ((n * rows * cols) // 4)
    plane = (m + 7) // 8 * 8 + 64
    p = Guarded((3, plane), torch.bfloat16, 777.0)
    xf = x.reshape(-1)[:m].contiguous()
    L.check(L.lib().rgnn_tc_split3(L.ptr(xf), m, plane, L.ptr(p.view), L.stream()), "split3")
    p.check("split3")


@pytest.mark.parametrize("m,n,k", [(77, 128, 64), (300, 256, 128)])
def test_tc3_gemm_guards(R, m, n, k):
    """fp32 outputs of the split-plane GEMMs leave through the 32-column TMA-store slabs: rows beyond M must be clipped."""
    L = R._lib
    a, b = torch.randn((m, k), device="cuda"), torch.randn((n, k), device="cuda") * 0.1
    prow = (m + 63) // 64 * 64
    a3 = torch.empty((3, prow * k), dtype=torch.bfloat16, device="cuda")
    L.check(L.lib().rgnn_tc_split3(L.ptr(a), m * k, prow * k, L.ptr(a3), L.stream()), "split3")
    b3 = torch.empty((3, n, k), dtype=torch.bfloat16, device="cuda")
    L.check(L.lib().rgnn_tc_split3_weights(L.ptr(b), n, k, L.ptr(b3), None, L.stream()), "split3w")
    c = Guarded((m, n), torch.float32, 9e9)
    L.check(L.lib().rgnn_tc3_gemm_tn(L.ptr(a3), k, prow, L.ptr(b3), k, n, L.ptr(c.view), n, m, n, k, 0, None, 1.0, L.stream()), "tn3")
    c.check("tc3_gemm_tn")
    ref = a.double() @ b.double().t()
    assert float((c.view.double() - ref).abs().max()) <= 1e-5 * float(ref.abs().max())
