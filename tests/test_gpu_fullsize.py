"""GPU: BASELINE.json-sized runs checked through size-independent properties (the oracle cannot finish these sizes in
seconds): mass conservation of the mean aggregation, linearity of the block in h, the adjoint identity <dy, L h> = <L^T dy, h>
that ties backward to forward, and agreement between the tensor-core path and the CUDA-core path on a slice."""
import types

import pytest
import torch

from conftest import nerr

pytestmark = pytest.mark.gpu


def make(R, f, comp=True, multi=True, seed=1):
    from robot_check import make_block
    from oracle import gnn_oracle as O
    blk, _ = make_block(R, O, f, comp, multi, seed=seed)
    return blk.cuda()


@pytest.mark.parametrize("G,K,F,dtype", [(4096, 5, 512, torch.bfloat16), (1024, 5, 1280, torch.float32), (512, 8, 512, torch.bfloat16)])
def test_mean_aggregation_conserves_mass(R, G, K, F, dtype):
    """copy_u-mean over complete graphs: every robot feeds K-1 destinations with weight 1/(K-1) => per-graph sums are preserved."""
    g = R.robot_batch(G, K, device="cuda")
    gen = torch.Generator(device="cuda").manual_seed(0)
    h = torch.rand((G * K, F, 8, 8), generator=gen, device="cuda").to(dtype)
    agg = R.copy_u_reduce(g, h, "mean")
    s_in = h.float().view(G, K, -1).sum(1)
    s_out = agg.float().view(G, K, -1).sum(1)
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert nerr(s_out, s_in) <= tol
    # FiLM with gamma = 1, beta = 0 is the same operator
    z = torch.zeros((g.num_edges(), 2 * F), device="cuda", dtype=dtype)
    z[:, 0::2] = 1
    assert torch.equal(R.film_mean(g, h, z), agg) or nerr(R.film_mean(g, h, z).float(), agg.float()) <= tol


@pytest.mark.parametrize("G,K,F", [(4096, 5, 512), (256, 5, 2048)])
def test_block_linearity_and_adjoint_bf16(R, G, K, F):
    """With the pose (hence gamma/beta) fixed the block is affine in h.  Checks y(a)+y(b)-y(0) = y(a+b) and
    <dy, y(h) - y(0)> = <dh, h> at the bench size (C2) and at ResNet-50 width (C4), bf16 tensor-core path."""
    blk = make(R, F)
    g = R.robot_batch(G, K, device="cuda")
    gen = torch.Generator(device="cuda").manual_seed(3)
    N, E = G * K, G * K * (K - 1)
    pose = torch.randn((E, 9), generator=gen, device="cuda")
    a = torch.rand((N, F, 8, 8), generator=gen, device="cuda").bfloat16()
    b = torch.rand((N, F, 8, 8), generator=gen, device="cuda").bfloat16()
    with torch.no_grad():
        y0 = blk(g, torch.zeros_like(a), pose).float()
        ya, yb = blk(g, a, pose).float(), blk(g, b, pose).float()
        yab = blk(g, (a.float() + b.float()).bfloat16(), pose).float()
    assert nerr(ya + yb - y0, yab) <= 2e-2
    dy = (ya - y0).bfloat16()          # dy = L a makes both sides ||L a||^2: no cancellation, so rounding noise stays relative
    h = a.clone().requires_grad_()
    y = blk(g, h, pose)
    y.backward(dy)
    lhs = ((y.detach().float() - y0) * dy.float()).sum().item()
    rhs = (h.grad.float() * a.float()).sum().item()
    assert abs(lhs - rhs) <= 2e-2 * max(abs(lhs), abs(rhs))
    for p in blk.parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all()


def test_tensor_core_path_equals_cuda_core_path_on_slice(R):
    """C4 width (F=2048): the first graphs of a tensor-core run must match a CUDA-core (rgnn_gemm) run of that slice."""
    F, K = 2048, 5
    blk = make(R, F, multi=False)
    gen = torch.Generator(device="cuda").manual_seed(8)
    G, Gs = 64, 4
    pose = torch.randn((G * K * (K - 1), 9), generator=gen, device="cuda")
    h = torch.rand((G * K, F, 8, 8), generator=gen, device="cuda").bfloat16()
    with torch.no_grad():
        y = blk(R.robot_batch(G, K, device="cuda"), h, pose)
        R._lib.USE_TC = False
        try:
            ys = blk(R.robot_batch(Gs, K, device="cuda"), h[: Gs * K].contiguous(), pose[: Gs * K * (K - 1)].contiguous())
        finally:
            R._lib.USE_TC = True
    assert nerr(y[: Gs * K].float(), ys.float()) <= 2e-2


def test_edge_cases_small_batches(R):
    """G = 1, a single isolated robot inside a batch, and P != 64 all run (general kernels) and keep zero rows for isolated nodes."""
    from oracle import gnn_oracle as O
    import numpy as np
    ks = [1, 5, 2]
    ed = [O.fully_connected_edges(k) for k in ks]
    g = R.RobotGraph(torch.from_numpy(np.concatenate([e[0] for e in ed])).cuda(), torch.from_numpy(np.concatenate([e[1] for e in ed])).cuda(),
                     ks, [len(e[0]) for e in ed])
    h = torch.rand((8, 16, 3, 3), device="cuda")
    z = torch.rand((g.num_edges(), 32), device="cuda")
    agg = R.film_mean(g, h, z)
    assert agg[0].abs().max().item() == 0.0       # isolated robot: DGL leaves zeros
    b = O.batch_graphs([e[0] for e in ed], [e[1] for e in ed], ks)
    assert torch.equal(agg.cpu(), O.film_mean(h.cpu(), z.cpu(), b["src"], b["dst"]))
    g1 = R.robot_batch(1, 5, device="cuda")
    h1 = torch.rand((5, 64, 8, 8), device="cuda").bfloat16()
    z1 = torch.rand((20, 128), device="cuda").bfloat16()
    ref = O.film_mean(h1.float().cpu(), z1.float().cpu(), *O.fully_connected_edges(5))
    assert nerr(R.film_mean(g1, h1, z1).float(), ref) <= 2e-2
