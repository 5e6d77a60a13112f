"""GPU: the tcgen05 (tensor-core) transforms, called straight through the C ABI, against an fp32 PyTorch reference of
the same bf16 operands, and against the CUDA-core rgnn_gemm path on whole layers.  bf16 tolerance 2e-2 (north_star);
the unit products here are far tighter because both sides see identical bf16 inputs (only accumulation order differs)."""
import types

import pytest
import torch

from conftest import nerr

pytestmark = pytest.mark.gpu


def bf(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(torch.bfloat16)


@pytest.mark.parametrize("m,n,k", [(128, 128, 64), (300, 256, 128), (1000, 1024, 512), (77, 384, 192)])
@pytest.mark.parametrize("epi", ["plain", "bias_sigmoid", "relu_mask"])
def test_tc_gemm_tn(R, m, n, k, epi):
    L = R._lib
    a, b = bf(m, k, seed=1), bf(n, k, seed=2, scale=0.1)
    c = torch.zeros((m, n), dtype=torch.bfloat16, device="cuda")
    bias = torch.randn(n, device="cuda")
    aux = bf(m, n, seed=3)
    ref = a.float() @ b.float().t()
    if epi == "plain":
        flags, pb, pa = 0, None, None
    elif epi == "bias_sigmoid":
        flags, pb, pa = L.EPI_BIAS_N | L.EPI_SIGMOID, bias, None
        ref = torch.sigmoid(ref + bias)
    else:
        flags, pb, pa = L.EPI_RELU_MASK, None, aux
        ref = ref * (aux.float() > 0)
    L.check(L.lib().rgnn_tc_gemm_tn(L.ptr(a), k, L.ptr(b), k, L.ptr(c), n, m, n, k, flags, L.ptr(pb), L.ptr(pa), 1.0, 0.0, L.stream()), "tn")
    torch.cuda.synchronize()
    assert nerr(c.float(), ref) <= 6e-3      # bf16 output rounding = 2^-8 relative


@pytest.mark.parametrize("m,n,k,splits", [(128, 128, 64, 1), (256, 256, 1000, 3), (1024, 512, 4096, 9), (128, 384, 130, 4)])
def test_tc_gemm_nn(R, m, n, k, splits):
    L = R._lib
    a, b = bf(k, m, seed=4), bf(k, n, seed=5)
    c = torch.full((m, n), 7.0, device="cuda")
    part = torch.empty((splits, m, n), device="cuda")
    L.check(L.lib().rgnn_tc_gemm_nn(L.ptr(a), m, L.ptr(b), n, L.ptr(c), n, m, n, k, 1.0, L.ptr(part), splits, L.stream()), "nn")
    torch.cuda.synchronize()
    assert nerr(c, a.float().t() @ b.float()) <= 1e-5


@pytest.mark.parametrize("nodes,c0,c1,m0,m1,bias", [(4, 64, 0, 128, 0, False), (10, 128, 128, 128, 0, True), (7, 256, 0, 128, 256, False),
                                                     (41, 512, 512, 512, 0, True)])
def test_tc_conv1x1(R, nodes, c0, c1, m0, m1, bias):
    L = R._lib
    w = bf(m0 + m1, c0 + c1, seed=6, scale=0.05)
    x0, x1 = bf(nodes, c0, 64, seed=7), (bf(nodes, c1, 64, seed=8) if c1 else None)
    o0 = torch.zeros((nodes, m0, 64), dtype=torch.bfloat16, device="cuda")
    o1 = torch.zeros((nodes, m1, 64), dtype=torch.bfloat16, device="cuda") if m1 else None
    b = torch.randn(m0 + m1, device="cuda") if bias else None
    L.check(L.lib().rgnn_tc_conv1x1(L.ptr(w), L.ptr(x0), c0, L.ptr(x1), c1, L.ptr(o0), m0, L.ptr(o1), m1, L.ptr(b), nodes, L.stream()), "conv")
    torch.cuda.synchronize()
    x = torch.cat([x0, x1], 1) if c1 else x0
    ref = torch.einsum("mk,vkp->vmp", w.float(), x.float())
    if bias:
        ref = ref + b.view(1, -1, 1)
    got = torch.cat([o0, o1], 1) if m1 else o0
    assert nerr(got.float(), ref) <= 6e-3


@pytest.mark.parametrize("nodes,m,c0,c1,splits", [(3, 128, 256, 0, 1), (20, 256, 256, 256, 4), (203, 512, 512, 512, 9)])
def test_tc_conv1x1_wgrad(R, nodes, m, c0, c1, splits):
    L = R._lib
    dy = bf(nodes, m, 64, seed=9)
    x0, x1 = bf(nodes, c0, 64, seed=10), (bf(nodes, c1, 64, seed=11) if c1 else None)
    dw = torch.full((m, c0 + c1), 3.0, device="cuda")
    db = torch.full((m,), 5.0, device="cuda")
    part = torch.empty((splits * m * (c0 + c1 + 1),), device="cuda")
    for with_bias in (True, False):
        L.check(L.lib().rgnn_tc_conv1x1_wgrad(L.ptr(dy), m, L.ptr(x0), c0, L.ptr(x1), c1, L.ptr(dw), c0 + c1,
                                              L.ptr(db) if with_bias else None, L.ptr(part), splits, nodes, L.stream()), "wgrad")
        torch.cuda.synchronize()
        x = torch.cat([x0, x1], 1) if c1 else x0
        assert nerr(dw, torch.einsum("vmp,vkp->mk", dy.float(), x.float())) <= 1e-5
    assert nerr(db, dy.float().sum((0, 2))) <= 1e-5


def test_cast_weights(R):
    L = R._lib
    w = torch.randn((200, 72), device="cuda")
    o = torch.empty((200, 72), dtype=torch.bfloat16, device="cuda")
    ot = torch.empty((72, 200), dtype=torch.bfloat16, device="cuda")
    L.check(L.lib().rgnn_tc_cast_weights(L.ptr(w), 200, 72, L.ptr(o), L.ptr(ot), L.stream()), "cast")
    assert torch.equal(o, w.to(torch.bfloat16)) and torch.equal(ot, w.t().contiguous().to(torch.bfloat16))


@pytest.mark.parametrize("g_,k,f", [(6, 5, 256), (8, 5, 512), (3, 4, 384), (2, 7, 640), (1, 2, 128)])
def test_block_tc_matches_cuda_core_path_and_oracle(R, g_, k, f):
    """Whole 2-layer block in bf16: tensor-core path vs the CUDA-core rgnn_gemm path (same bf16 operands) and vs the fp64 oracle."""
    from robot_check import check_block, make_block
    from oracle import gnn_oracle as O
    errs = check_block(R, O, g_, k, f, torch.bfloat16, 2e-2)
    assert max(errs.values()) <= 4e-2
    sb = O.synthetic_batch(g_, k, f, p=64, seed=3)
    outs = []
    for use_tc in (True, False):
        R._lib.USE_TC = use_tc
        try:
            blk, _ = make_block(R, O, f, True, True)
            blk.cuda()
            g = R.robot_batch(g_, k, device="cuda")
            h = sb["h"].reshape(g_ * k, f, 8, 8).cuda().bfloat16().requires_grad_()
            y = blk(g, h, sb["pose"].cuda())
            y.backward(sb["dy"].reshape(y.shape).cuda().bfloat16())
            outs.append((y.float(), h.grad.float(), blk.conv1.weight.grad.clone(), blk.gcn1.edge_encoder.layers[2].weight.grad.clone()))
        finally:
            R._lib.USE_TC = True
    for a, b in zip(*outs):
        assert nerr(a, b) <= 2e-2


@pytest.mark.parametrize("m,n,k", [(200, 256, 256), (77, 128, 128)])
def test_tc_gemm_tn_axpby_identity_mapping(R, m, n, k):
    """alpha * A.B^T + beta * aux + bias: GCN2Conv's (1-b) X + b X W in one epilogue."""
    L = R._lib
    a, b = bf(m, k, seed=21), bf(n, k, seed=22, scale=0.1)
    aux = bf(m, n, seed=23)
    bias = torch.randn(n, device="cuda")
    c = torch.zeros((m, n), dtype=torch.bfloat16, device="cuda")
    L.check(L.lib().rgnn_tc_gemm_tn(L.ptr(a), k, L.ptr(b), k, L.ptr(c), n, m, n, k, L.EPI_ADD_AUX | L.EPI_BIAS_N, L.ptr(bias), L.ptr(aux),
                                    0.3, 0.7, L.stream()), "tn")
    torch.cuda.synchronize()
    assert nerr(c.float(), 0.3 * (a.float() @ b.float().t()) + 0.7 * aux.float() + bias) <= 6e-3


@pytest.mark.parametrize("which", ["graphconv", "gcn2"])
def test_dgl_style_transforms_on_tensor_cores_bf16(R, which):
    """GraphConv / GCN2Conv (north-star A11) in bf16 with D = 256: weight transform on tcgen05, vs the fp64 DGL formula."""
    from oracle import gnn_oracle as O
    gen = torch.Generator().manual_seed(6)
    g_, k, d = 40, 5, 256
    sb = O.synthetic_batch(g_, k, 8, p=1, seed=6)
    n = g_ * k
    x, x0 = torch.randn((n, d), generator=gen), torch.randn((n, d), generator=gen)
    w = torch.randn((d, d), generator=gen) * 0.05
    b = torch.randn((d,), generator=gen)
    dy = torch.randn((n, d), generator=gen)
    t = [v.double().requires_grad_() for v in (x, x0, w, b)]
    g = R.robot_batch(g_, k, device="cuda")
    xc, x0c = x.cuda().bfloat16().requires_grad_(), x0.cuda().bfloat16().requires_grad_()
    if which == "graphconv":
        ref = O.graphconv(t[0], sb["src"], sb["dst"], t[2], t[3], "both")
        conv = R.GraphConv(d, d, norm="both")
        conv.load_state_dict({"weight": w, "bias": b})
        y = conv.cuda()(g, xc)
        wparam = conv.weight
    else:
        ref = O.gcn2conv(t[0], t[1], sb["src"], sb["dst"], t[2], t[3], layer=2, alpha=0.1, lambda_=1.0)
        conv = R.GCN2Conv(d, layer=2, alpha=0.1, lambda_=1)
        conv.load_state_dict({"weight1": w, "bias": b})
        y = conv.cuda()(g, xc, x0c)
        wparam = conv.weight1
    ref.backward(dy.double())
    y.backward(dy.cuda().bfloat16())
    assert y.dtype == torch.bfloat16
    assert nerr(y.float(), ref) <= 2e-2
    assert nerr(xc.grad.float(), t[0].grad) <= 2e-2
    assert nerr(wparam.grad, t[2].grad) <= 4e-2 and nerr(conv.bias.grad, t[3].grad) <= 4e-2
    if which == "gcn2":
        assert nerr(x0c.grad.float(), t[1].grad) <= 2e-2


@pytest.mark.parametrize("k,gN,f,fo", [(5, 6, 64, 128), (5, 40, 512, 512), (3, 7, 96, 128), (8, 5, 128, 256), (2, 9, 32, 128), (6, 4, 256, 128)])
def test_fused_film_conv_forward(R, k, gN, f, fo):
    """rgnn_tc_conv_film_fused vs the two separate kernels: the aggregated features it writes out are BIT-identical to
    k_film_fwd_complete (same fp32 formula, same order), y agrees to bf16 rounding (the K loop interleaves the halves)."""
    L = R._lib
    g = R.robot_batch(gN, k, device="cuda")
    ix = g.index()
    h = bf(gN * k, f, 64, seed=31).abs()
    z = torch.rand((ix.E, 2 * f), generator=torch.Generator(device="cuda").manual_seed(32), device="cuda").bfloat16()
    w = bf(fo, 2 * f, seed=33, scale=0.05)
    bias = torch.randn(fo, device="cuda")
    gh_ref = torch.empty_like(h)
    R.functional._aggregate(ix, h, z, True, None, L.POST_MEAN, None, None, gh_ref, None, False)
    y_ref = torch.empty((gN * k, fo, 64), dtype=torch.bfloat16, device="cuda")
    if f % 64 == 0:
        L.check(L.lib().rgnn_tc_conv1x1(L.ptr(w), L.ptr(h), f, L.ptr(gh_ref), f, L.ptr(y_ref), fo, None, 0, L.ptr(bias), gN * k, L.stream()), "conv")
        y_ref = y_ref.float()
    else:
        y_ref = torch.einsum("ok,vkp->vop", w.float(), torch.cat([h, gh_ref], 1).float()) + bias.view(1, -1, 1)
    for store in (True, False):
        y = torch.zeros((gN * k, fo, 64), dtype=torch.bfloat16, device="cuda")
        gh = torch.zeros_like(h) if store else None
        L.check(L.lib().rgnn_tc_conv_film_fused(L.ptr(w), L.ptr(h), L.ptr(z), L.ptr(bias), L.ptr(y), L.ptr(gh), gN, k, f, fo, L.stream()), "fused")
        torch.cuda.synchronize()
        if store:
            assert torch.equal(gh, gh_ref)
        assert nerr(y.float(), y_ref) <= 6e-3


@pytest.mark.parametrize("train", [True, False])
def test_block_with_fused_prologue_matches_unfused(R, train):
    """The opt-in fused path (USE_FUSED) inside the whole block: same outputs / gradients as the default path; in eval the
    aggregated features are never written."""
    from robot_check import make_block
    from oracle import gnn_oracle as O
    G, K, F = 6, 5, 256
    sb = O.synthetic_batch(G, K, F, p=64, seed=4)
    res = []
    for fused in (False, True):
        R._lib.USE_FUSED = fused
        try:
            blk, _ = make_block(R, O, F, True, True)
            blk.cuda()
            g = R.robot_batch(G, K, device="cuda")
            h = sb["h"].reshape(G * K, F, 8, 8).cuda().bfloat16().requires_grad_(train)
            if train:
                y = blk(g, h, sb["pose"].cuda())
                y.backward(sb["dy"].reshape(y.shape).cuda().bfloat16())
                res.append((y.detach().float(), h.grad.float(), blk.conv1.weight.grad.clone(), blk.gcn2.edge_encoder.layers[2].weight.grad.clone()))
            else:
                with torch.no_grad():
                    res.append((blk(g, h, sb["pose"].cuda()).float(),))
        finally:
            R._lib.USE_FUSED = False
    for a, b in zip(*res):
        assert nerr(b, a) <= 1e-2


@pytest.mark.parametrize("g_,k,f", [(5, 5, 512), (3, 2, 64), (4, 3, 128), (7, 4, 256), (2, 6, 192), (150, 5, 128)])
def test_fused_dx_film_backward(R, g_, k, f):
    """north_star (c)/(d): the FiLM backward in the epilogue of the dX GEMM (rgnn_tc_conv_dx_film_bwd, default for bf16 complete graphs
    of 2..6 robots) against the unfused dX GEMM + FiLM backward kernel pair and against the fp64 oracle - dh and every gradient."""
    from robot_check import check_block, make_block
    from oracle import gnn_oracle as O
    L = R._lib
    assert L.USE_FUSED_BWD and f <= L.FUSED_BWD_MAX_F
    errs = check_block(R, O, g_, k, f, torch.bfloat16, 2e-2)
    assert max(errs.values()) <= 4e-2, errs
    sb = O.synthetic_batch(g_, k, f, p=64, seed=4)
    n = g_ * k
    dy = torch.randn((n, f, 8, 8), generator=torch.Generator().manual_seed(7)).cuda().bfloat16()
    res = []
    for fused in (True, False):
        L.USE_FUSED_BWD = fused
        try:
            blk, _ = make_block(R, O, f, True, True)
            blk.cuda()
            g = R.robot_batch(g_, k, device="cuda")
            h = sb["h"].reshape(n, f, 8, 8).cuda().bfloat16().requires_grad_()
            n0 = L.lib().rgnn_launch_count()
            y = blk(g, h, sb["pose"].cuda())
            y.backward(dy)
            torch.cuda.synchronize()
            res.append((y.detach().float(), h.grad.float(), {nm: p.grad.clone() for nm, p in blk.named_parameters()}, L.lib().rgnn_launch_count() - n0))
        finally:
            L.USE_FUSED_BWD = True
    (y1, dh1, g1, l1), (y0, dh0, g0, l0) = res
    assert torch.equal(y1, y0)                       # same forward
    # per layer, when the fused kernel applies (the tcgen05 conv needs Fo % 128 == 0): the FiLM backward launch is gone and so is the
    # edge encoder's second bias-gradient reduction (column sums of dz inside the fused kernel)
    assert l1 == (l0 - 4 if f % 128 == 0 else l0)
    assert nerr(dh1, dh0) <= 2e-2                    # dcat is rounded to bf16 only in the unfused path
    for nm in g1:
        assert nerr(g1[nm], g0[nm]) <= 2e-2, nm


def test_fused_dx_film_backward_wide_layer_when_forced(R):
    """Above FUSED_BWD_MAX_F the unfused pair is the default (it is faster there); the fused kernel itself is still correct at
    ResNet-50 width (Fo = 2048: 32 k-blocks per half-tile) when forced."""
    from robot_check import check_block
    from oracle import gnn_oracle as O
    L = R._lib
    old, L.FUSED_BWD_MAX_F = L.FUSED_BWD_MAX_F, 1 << 20
    try:
        n0 = L.lib().rgnn_launch_count()
        errs = check_block(R, O, 2, 5, 2048, torch.bfloat16, 2e-2, multi=False)
        n_forced = L.lib().rgnn_launch_count() - n0
    finally:
        L.FUSED_BWD_MAX_F = old
    assert max(errs.values()) <= 4e-2, errs
    n0 = L.lib().rgnn_launch_count()
    check_block(R, O, 2, 5, 2048, torch.bfloat16, 2e-2, multi=False)
    assert L.lib().rgnn_launch_count() - n0 == n_forced + 2      # FiLM backward + the bias-gradient reduction
