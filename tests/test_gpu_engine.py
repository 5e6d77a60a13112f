"""GPU: the host-buffer pipeline (chunked, three streams) must give the same outputs and parameter gradients as one
device-resident fwd+bwd over the whole batch."""
import pytest
import torch

from conftest import nerr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
def test_host_pipeline_matches_single_shot(R, dtype, tol):
    from robot_check import make_block
    from oracle import gnn_oracle as O
    G, K, F = 24, 5, 128
    blk, _ = make_block(R, O, F, True, True)
    blk.cuda()
    sb = O.synthetic_batch(G, K, F, p=64, seed=5)
    N = G * K
    h = sb["h"].reshape(N, F, 8, 8).to(dtype)
    dy = sb["dy"].reshape(N, F, 8, 8).to(dtype)
    pose = sb["pose"]
    # single shot
    g = R.robot_batch(G, K, device="cuda")
    hd = h.cuda().requires_grad_()
    y = blk(g, hd, pose.cuda())
    y.backward(dy.cuda())
    ref = {n: p.grad.clone() for n, p in blk.named_parameters()}
    y_ref, dh_ref = y.detach().float().cpu(), hd.grad.float().cpu()
    # pipelined from pinned host memory
    pipe = R.HostPipeline(blk, G, K, F, dtype=dtype, device="cuda", chunks=4, slots=2)
    hh, hdy, hp = h.pin_memory(), dy.pin_memory(), pose.pin_memory()
    oy, odh = torch.empty_like(hh).pin_memory(), torch.empty_like(hh).pin_memory()
    for _ in range(2):      # second call reuses the slots
        pipe.step(hh, hp, hdy, oy, odh)
        torch.cuda.synchronize()
    assert nerr(oy.float(), y_ref) <= tol and nerr(odh.float(), dh_ref) <= tol
    for n, p in blk.named_parameters():
        assert nerr(p.grad, ref[n]) <= 2 * tol, n
    # streaming mode: three steps back to back without joining the calling stream, then one wait()
    oy.zero_(); odh.zero_()
    for _ in range(3):
        pipe.step(hh, hp, hdy, oy, odh, wait=False)
    pipe.wait()
    torch.cuda.current_stream().synchronize()
    assert nerr(oy.float(), y_ref) <= tol and nerr(odh.float(), dh_ref) <= tol
    for n, p in blk.named_parameters():
        assert nerr(p.grad, ref[n]) <= 2 * tol, n


@pytest.mark.parametrize("dtype,f", [(torch.bfloat16, 512), (torch.float32, 64)])
def test_cuda_graph_replay_matches_eager(R, dtype, f):
    """Script shape (G=8, K=5): the captured step must reproduce the eager step bit for bit, for fresh inputs on every replay."""
    from robot_check import make_block
    from oracle import gnn_oracle as O
    G, K = 8, 5
    blk, _ = make_block(R, O, f, True, True)
    blk.cuda()
    g = R.robot_batch(G, K, device="cuda")
    gen = torch.Generator(device="cuda").manual_seed(0)
    mk = lambda: (torch.rand((G * K, f, 8, 8), generator=gen, device="cuda").to(dtype),
                  torch.randn((G * K * (K - 1), 9), generator=gen, device="cuda"),
                  torch.randn((G * K, f, 8, 8), generator=gen, device="cuda").to(dtype))
    h0, p0, d0 = mk()
    step = R.GraphedStep(blk, g, h0, p0, d0)
    for _ in range(3):
        h, pose, dy = mk()
        y, dh = step(h, pose, dy)
        torch.cuda.synchronize()
        y, dh = y.clone(), dh.clone()
        grads = {n: p.grad.clone() for n, p in blk.named_parameters()}
        he = h.clone().requires_grad_()
        for p in blk.parameters():
            p.grad = None
        ye = blk(g, he, pose)
        ye.backward(dy)
        assert torch.equal(y, ye) and torch.equal(dh, he.grad)
        for n, p in blk.named_parameters():
            # atomics in the bias / edge-MLP layer-1 gradient kernels reorder fp32 sums between runs
            assert nerr(grads[n], p.grad) <= 1e-5, n
