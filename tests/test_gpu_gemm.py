"""rgnn_gemm (CUDA-core fp32 GEMM family, the parity path's products) against fp64 torch on the same operands: both kernels behind
the entry (the 64x64 general one and the 128-wide FFMA2 one that takes the large all-fp32 products), every layout the layers use
(dgl/model/models.py:146-151 linear layers, :184 1x1 conv per node, their autograd), ragged edges, folded node pairs, split-K."""
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 1e-5          # north_star fp32 tolerance (normalised max error)


@pytest.fixture(scope="module")
def R():
    import robot_gnn_b200 as R
    return R


def nerr(a, ref):
    return float((a.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("m,n,k", [(64, 64, 2500), (200, 136, 300), (1280, 2560, 96), (129, 65, 1000), (4096, 64, 80)])
@pytest.mark.parametrize("epi", ["none", "bias_n_sigmoid", "relu_mask"])
def test_linear_layer_products(R, m, n, k, epi):
    """y[m, n] = x[m, k] . w[n, k]^T (+ bias, sigmoid | ReLU mask) - the edge encoder's layers and their dX."""
    F, L = R.functional, R._lib
    g = torch.Generator(device="cuda").manual_seed(m + n + k)
    x = torch.randn((m, k), generator=g, device="cuda")
    w = torch.randn((n, k), generator=g, device="cuda") / k ** 0.5
    b = torch.randn((n,), generator=g, device="cuda")
    aux = torch.randn((m, n), generator=g, device="cuda")
    y = torch.empty((m, n), device="cuda")
    ref = x.double() @ w.double().t()
    if epi == "bias_n_sigmoid":
        F._gemm(x, (0, 0, k, 1), w, (0, 0, 1, k), y, (0, n, 1), m, n, k, epi=L.EPI_BIAS_N | L.EPI_SIGMOID, bias=b)
        ref = torch.sigmoid(ref + b.double())
    elif epi == "relu_mask":
        F._gemm(x, (0, 0, k, 1), w, (0, 0, 1, k), y, (0, n, 1), m, n, k, epi=L.EPI_RELU_MASK, aux=aux)
        ref = ref * (aux > 0)
    else:
        F._gemm(x, (0, 0, k, 1), w, (0, 0, 1, k), y, (0, n, 1), m, n, k)
    assert nerr(y, ref) <= TOL


@pytest.mark.parametrize("nodes,fo,f", [(7, 128, 256), (10, 200, 64), (1, 64, 512), (33, 1280, 160)])
def test_per_node_conv_products(R, nodes, fo, f):
    """y_v[fo, 64] = W[fo, f] . x_v[f, 64] + b per node (odd node counts exercise the folded last tile), then the in-place
    accumulate form (ADD_AUX on the output itself) the concat's second half uses."""
    F, L = R.functional, R._lib
    p = 64
    g = torch.Generator(device="cuda").manual_seed(nodes + fo)
    w = torch.randn((fo, 2 * f), generator=g, device="cuda") / f ** 0.5
    x = torch.randn((nodes, f, p), generator=g, device="cuda")
    x2 = torch.randn((nodes, f, p), generator=g, device="cuda")
    b = torch.randn((fo,), generator=g, device="cuda")
    y = torch.empty((nodes, fo, p), device="cuda")
    sb, sc = (f * p, 0, p, 1), (fo * p, p, 1)
    F._gemm(w, (0, 0, 2 * f, 1), x, sb, y, sc, fo, p, f, batch=nodes, epi=L.EPI_BIAS_M, bias=b)
    ref = torch.einsum("ok,vkp->vop", w[:, :f].double(), x.double()) + b.double()[None, :, None]
    assert nerr(y, ref) <= TOL
    F._gemm(w[:, f:], (0, 0, 2 * f, 1), x2, sb, y, sc, fo, p, f, batch=nodes, epi=L.EPI_ADD_AUX, aux=y, beta=1.0)
    ref = ref + torch.einsum("ok,vkp->vop", w[:, f:].double(), x2.double())
    assert nerr(y, ref) <= TOL
    # dX_v = W^T . dY_v  (A read m-fast)
    dx = torch.empty((nodes, f, p), device="cuda")
    F._gemm(w, (0, 0, 1, 2 * f), y, (fo * p, 0, p, 1), dx, (f * p, p, 1), f, p, fo, batch=nodes)
    assert nerr(dx, torch.einsum("ok,vop->vkp", w[:, :f].double(), y.double())) <= TOL


@pytest.mark.parametrize("nodes,fo,f,split", [(40, 128, 192, 1), (300, 256, 128, 7), (64, 70, 130, 64)])
def test_weight_gradient_products(R, nodes, fo, f, split):
    """dW[o, k] = sum_v sum_p dy[v, o, p] x[v, k, p]: K = 64 per node, reduced over the nodes, optionally split."""
    F = R.functional
    p = 64
    g = torch.Generator(device="cuda").manual_seed(nodes)
    dy = torch.randn((nodes, fo, p), generator=g, device="cuda")
    x = torch.randn((nodes, f, p), generator=g, device="cuda")
    dw = torch.empty((fo, f), device="cuda")
    partial = torch.empty((split, fo, f), device="cuda") if split > 1 else None
    F._gemm(dy, (0, fo * p, p, 1), x, (0, f * p, 1, p), dw, (0, f, 1), fo, f, p, red=nodes, split=split, partial=partial)
    ref = torch.einsum("vop,vkp->ok", dy.double(), x.double())
    assert nerr(dw, ref) <= TOL
    # linear-layer weight gradient: dW[n, k] = sum_rows dy[r, n] x[r, k]  (both operands read along their slow dimension)
    rows = nodes * 16
    a = torch.randn((rows, fo), generator=g, device="cuda")
    b = torch.randn((rows, f), generator=g, device="cuda")
    out = torch.empty((fo, f), device="cuda")
    sp = max(1, min(split, (rows + 15) // 16))
    part = torch.empty((sp, fo, f), device="cuda") if sp > 1 else None
    F._gemm(a, (0, 0, 1, fo), b, (0, 0, f, 1), out, (0, f, 1), fo, f, rows, split=sp, partial=part)
    assert nerr(out, a.double().t() @ b.double()) <= TOL


@pytest.mark.parametrize("k", [2, 5, 8])
@pytest.mark.parametrize("dtype,row,tol", [(torch.float32, 2048, 1e-5), (torch.bfloat16, 4096, 2e-2), (torch.float32, 1000, 1e-5)])
def test_mean_mix_matches_dense_formula(R, k, dtype, row, tol):
    """GCN2Conv propagation + initial residual on complete graphs, r = sa * A_hat x + sb * x0 with A_hat = D^-1/2 A D^-1/2 (SURVEY
    App. B), forward and both gradients, against the dense fp64 formula; row = 1000 takes the aggregate + axpby fallback."""
    g_ = 6
    gen = torch.Generator(device="cuda").manual_seed(k + row)
    x = torch.randn((g_ * k, row), generator=gen, device="cuda").to(dtype).requires_grad_()
    x0 = torch.randn((g_ * k, row), generator=gen, device="cuda").to(dtype).requires_grad_()
    dy = torch.randn((g_ * k, row), generator=gen, device="cuda").to(dtype)
    graph = R.robot_batch(g_, k, device="cuda")
    y = R.functional.mean_mix(graph, x, x0, 0.9, 0.1)
    y.backward(dy)
    a_hat = (torch.ones(k, k, dtype=torch.float64, device="cuda") - torch.eye(k, dtype=torch.float64, device="cuda")) / (k - 1)
    xd = x.detach().double().reshape(g_, k, row)
    ref = 0.9 * torch.einsum("ab,gbr->gar", a_hat, xd).reshape(g_ * k, row) + 0.1 * x0.detach().double()
    assert nerr(y.detach(), ref) <= tol
    dref = 0.9 * torch.einsum("ab,gbr->gar", a_hat, dy.double().reshape(g_, k, row)).reshape(g_ * k, row)
    assert nerr(x.grad, dref) <= tol
    assert nerr(x0.grad, 0.1 * dy.double()) <= tol
