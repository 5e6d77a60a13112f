"""GPU: the fp32 PARITY mode on the tensor cores - split-bf16 products (rgnn_tc_split3*, rgnn_tc3_*), called straight through
the C ABI against fp64 references of the same fp32 operands.  Tolerance 1e-5 normalised max error (BASELINE.json north_star,
fp32); profiles/r02_fp32_split_study.jsonl predicts <= 2.5e-6 for reductions of up to 2560 elements per accumulator."""
import types

import pytest
import torch

from conftest import nerr

pytestmark = pytest.mark.gpu

TOL = 1e-5


def rnd(*shape, seed=0, scale=1.0, relu=False):
    g = torch.Generator(device="cuda").manual_seed(seed)
    t = torch.randn(shape, generator=g, device="cuda") * scale
    return torch.relu(t) if relu else t


def split3(L, x, n=None, plane=None):
    n = x.numel() if n is None else n
    plane = n if plane is None else plane
    out = torch.full((3, plane), 9.0, dtype=torch.bfloat16, device="cuda")
    L.check(L.lib().rgnn_tc_split3(L.ptr(x), n, plane, L.ptr(out), L.stream()), "split3")
    return out


def test_split3_is_an_exact_decomposition(R):
    L = R._lib
    x = rnd(1000, 64, seed=1) * torch.exp(rnd(1000, 64, seed=2) * 4)         # wide dynamic range
    x[0, :8] = torch.tensor([0.0, -0.0, 1.0, -1.0, 3.0e38, 1.0e-30, 65504.0, 1.0 + 2.0 ** -23], device="cuda")
    n, plane = x.numel(), x.numel() + 64
    p = split3(L, x, n, plane)
    s = p[0, :n].double() + p[1, :n].double() + p[2, :n].double()
    # three bf16 terms carry 24 significand bits: the sum reproduces fp32 to within its last bit
    assert float(((s - x.double().reshape(-1)).abs() / x.double().reshape(-1).abs().clamp(min=1e-300)).max()) <= 2.0 ** -23
    assert bool((p[:, n:] == 0).all())                                        # K tails are zero
    assert torch.equal(p[0, :n], x.reshape(-1).bfloat16())                    # plane 0 is the bf16 rounding of x
    # weights: planes and planes of the transpose
    w = rnd(200, 72, seed=3)
    o = torch.empty((3, 200, 72), dtype=torch.bfloat16, device="cuda")
    ot = torch.empty((3, 72, 200), dtype=torch.bfloat16, device="cuda")
    L.check(L.lib().rgnn_tc_split3_weights(L.ptr(w), 200, 72, L.ptr(o), L.ptr(ot), L.stream()), "split3w")
    assert torch.equal(o, split3(L, w.reshape(-1).contiguous()).view(3, 200, 72))
    assert torch.equal(ot, o.transpose(1, 2).contiguous())


@pytest.mark.parametrize("m,n,k", [(128, 128, 64), (300, 256, 128), (1000, 2560, 1280), (77, 384, 192), (4000, 1024, 512)])
@pytest.mark.parametrize("epi", ["plain", "bias_sigmoid"])
def test_tc3_gemm_tn(R, m, n, k, epi):
    """nn.Linear(F, 2F) + Sigmoid in fp32 (models.py:149-150): A = relu(.) activations, B = weights ~ U(+-1/sqrt(K))."""
    L = R._lib
    a = rnd(m, k, seed=1, relu=True)
    b = (torch.rand((n, k), generator=torch.Generator(device="cuda").manual_seed(2), device="cuda") * 2 - 1) / k ** 0.5
    prow = (m + 63) // 64 * 64
    a3 = split3(L, a, m * k, prow * k)
    b3 = torch.empty((3, n, k), dtype=torch.bfloat16, device="cuda")
    L.check(L.lib().rgnn_tc_split3_weights(L.ptr(b), n, k, L.ptr(b3), None, L.stream()), "split3w")
    c = torch.full((m, n), 7.0, device="cuda")
    bias = rnd(n, seed=3)
    ref = a.double() @ b.double().t()
    flags, pb = 0, None
    if epi == "bias_sigmoid":
        flags, pb = L.EPI_BIAS_N | L.EPI_SIGMOID, bias
        ref = torch.sigmoid(ref + bias.double())
    L.check(L.lib().rgnn_tc3_gemm_tn(L.ptr(a3), k, prow, L.ptr(b3), k, n, L.ptr(c), n, m, n, k, flags, L.ptr(pb), 1.0, L.stream()), "tn3")
    torch.cuda.synchronize()
    assert nerr(c, ref) <= TOL


@pytest.mark.parametrize("m,n,k,splits", [(128, 128, 64, 1), (256, 256, 1000, 3), (1024, 512, 20480, 9), (128, 384, 130, 4)])
def test_tc3_gemm_nn(R, m, n, k, splits):
    """dW2 = dZ^T . A1 in fp32: both operands MN-major, the edges are the reduction (split so no accumulator sums > 4096)."""
    L = R._lib
    a, b = rnd(k, m, seed=4), rnd(k, n, seed=5, relu=True)
    prow = (k + 63) // 64 * 64
    a3, b3 = split3(L, a, k * m, prow * m), split3(L, b, k * n, prow * n)
    splits = max(splits, (k + 4095) // 4096)
    c = torch.full((m, n), 7.0, device="cuda")
    part = torch.empty((splits, m, n), device="cuda")
    L.check(L.lib().rgnn_tc3_gemm_nn(L.ptr(a3), m, prow, L.ptr(b3), n, prow, L.ptr(c), n, m, n, k, 1.0, L.ptr(part), splits, L.stream()), "nn3")
    torch.cuda.synchronize()
    assert nerr(c, a.double().t() @ b.double()) <= TOL


@pytest.mark.parametrize("nodes,c0,c1,m0,m1,bias", [(4, 64, 0, 128, 0, False), (10, 128, 128, 128, 0, True), (7, 256, 0, 128, 256, False),
                                                     (41, 1280, 1280, 1280, 0, True), (9, 1280, 0, 1280, 1280, False)])
def test_tc3_conv1x1(R, nodes, c0, c1, m0, m1, bias):
    """Conv2d(2F, F, 1) on cat(h, g_h) in fp32 (models.py:164,182-184) and its dX (two outputs), incl. the C3 shape F = 1280."""
    L = R._lib
    k = c0 + c1
    w = (torch.rand((m0 + m1, k), generator=torch.Generator(device="cuda").manual_seed(6), device="cuda") * 2 - 1) / k ** 0.5
    x0, x1 = rnd(nodes, c0, 64, seed=7, relu=True), (rnd(nodes, c1, 64, seed=8, relu=True) if c1 else None)
    w3 = torch.empty((3, m0 + m1, k), dtype=torch.bfloat16, device="cuda")
    L.check(L.lib().rgnn_tc_split3_weights(L.ptr(w), m0 + m1, k, L.ptr(w3), None, L.stream()), "split3w")
    x03, x13 = split3(L, x0), (split3(L, x1) if c1 else None)
    o0 = torch.full((nodes, m0, 64), 7.0, device="cuda")
    o1 = torch.full((nodes, m1, 64), 7.0, device="cuda") if m1 else None
    b = rnd(m0 + m1, seed=9) if bias else None
    L.check(L.lib().rgnn_tc3_conv1x1(L.ptr(w3), L.ptr(x03), c0, L.ptr(x13), c1, nodes, L.ptr(o0), m0, L.ptr(o1), m1, L.ptr(b), nodes,
                                     L.stream()), "conv3")
    torch.cuda.synchronize()
    x = torch.cat([x0, x1], 1) if c1 else x0
    ref = torch.einsum("mk,vkp->vmp", w.double(), x.double())
    if bias:
        ref = ref + b.double().view(1, -1, 1)
    got = torch.cat([o0, o1], 1) if m1 else o0
    assert nerr(got, ref) <= TOL


@pytest.mark.parametrize("nodes,m,c0,c1,splits", [(3, 128, 256, 0, 1), (20, 256, 256, 256, 4), (203, 512, 512, 512, 9), (640, 128, 256, 256, 1)])
def test_tc3_conv1x1_wgrad(R, nodes, m, c0, c1, splits):
    L = R._lib
    dy = rnd(nodes, m, 64, seed=9)
    x0, x1 = rnd(nodes, c0, 64, seed=10, relu=True), (rnd(nodes, c1, 64, seed=11, relu=True) if c1 else None)
    dy3, x03, x13 = split3(L, dy), split3(L, x0), (split3(L, x1) if c1 else None)
    splits = max(splits, (nodes + 63) // 64)
    dw = torch.full((m, c0 + c1), 3.0, device="cuda")
    part = torch.empty((splits * m * (c0 + c1),), device="cuda")
    L.check(L.lib().rgnn_tc3_conv1x1_wgrad(L.ptr(dy3), m, L.ptr(x03), c0, L.ptr(x13), c1, nodes, L.ptr(dw), c0 + c1, L.ptr(part), splits, nodes,
                                           L.stream()), "wgrad3")
    torch.cuda.synchronize()
    x = torch.cat([x0, x1], 1) if c1 else x0
    assert nerr(dw, torch.einsum("vmp,vkp->mk", dy.double(), x.double())) <= TOL


@pytest.mark.parametrize("g_,k,f,comp,multi", [(8, 5, 1280, True, False), (4, 5, 512, True, True), (8, 5, 1280, False, False), (3, 4, 256, True, True)])
def test_fp32_block_on_tensor_cores_vs_oracle_and_cuda_cores(R, g_, k, f, comp, multi):
    """The whole fp32 block (C1 / C3 shapes among them) through the split-bf16 tensor-core products: within 1e-5 of the fp64 oracle,
    and the CUDA-core parity path (USE_TC3 off) stays available and agrees."""
    from robot_check import check_block
    from oracle import gnn_oracle as O
    L = R._lib
    n0 = L.lib().rgnn_launch_count()
    errs = check_block(R, O, g_, k, f, torch.float32, TOL, comp=comp, multi=multi)
    assert max(errs.values()) <= 2 * TOL, errs
    L.USE_TC3 = False
    try:
        errs_cc = check_block(R, O, g_, k, f, torch.float32, TOL, comp=comp, multi=multi)
    finally:
        L.USE_TC3 = True
    assert max(errs_cc.values()) <= 2 * TOL, errs_cc
    assert L.lib().rgnn_launch_count() > n0
