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


@pytest.mark.parametrize("dtype,f,tol", [(torch.bfloat16, 512, 2e-2), (torch.float32, 128, 1e-5)])
def test_weight_cast_cache_follows_the_weights(R, dtype, f, tol):
    """The bf16 / split-plane weight operands are cached on the parameters until they change: after an in-place update (what an
    optimizer step or load_state_dict does) forward and backward must use the NEW weights, and a loop over fixed weights must
    not launch cast kernels at all."""
    from robot_check import make_block, nerr
    from oracle import gnn_oracle as O
    G, K = 4, 5
    blk, _ = make_block(R, O, f, True, True)
    blk.cuda()
    g = R.robot_batch(G, K, device="cuda")
    g.index()                                                    # (the lazily built CSR index would count as launches of the first pass)
    sb = O.synthetic_batch(G, K, f, p=64, seed=2)
    h0 = sb["h"].reshape(G * K, f, 8, 8).cuda().to(dtype)
    pose, dy = sb["pose"].cuda(), sb["dy"].reshape(G * K, f, 8, 8).cuda().to(dtype)

    def run():
        h = h0.clone().requires_grad_()
        for p in blk.parameters():
            p.grad = None
        n0 = R._lib.lib().rgnn_launch_count()
        y = blk(g, h, pose)
        y.backward(dy)
        return y.detach().float(), h.grad.float(), R._lib.lib().rgnn_launch_count() - n0

    def oracle():
        ls = []
        for j in (1, 2):
            enc, conv = getattr(blk, f"gcn{j}").edge_encoder.layers, getattr(blk, f"conv{j}")
            ls.append({k: v.detach().double().cpu() for k, v in dict(w1=enc[0].weight, b1=enc[0].bias, w2=enc[2].weight, b2=enc[2].bias,
                                                                     wc=conv.weight.reshape(f, 2 * f), bc=conv.bias).items()})
        ho = h0.double().cpu().requires_grad_()
        yo = O.gnn_block(ho, pose.double().cpu(), sb["src"], sb["dst"], ls, compress_gcn=True)
        yo.backward(dy.double().cpu())
        return yo.detach(), ho.grad

    y1, dh1, l1 = run()
    y1b, dh1b, l1b = run()
    assert l1b < l1 and torch.equal(y1, y1b), (l1, l1b)         # second pass: no cast launches, same result
    yo, dho = oracle()
    assert nerr(y1, yo) <= tol and nerr(dh1, dho) <= tol
    with torch.no_grad():                                        # an "optimizer step"
        for p in blk.parameters():
            p.mul_(1.5)
    y2, dh2, l2 = run()
    assert l2 == l1, (l1, l2)                                    # the casts ran again
    yo2, dho2 = oracle()
    assert nerr(y2, yo2) <= tol and nerr(dh2, dho2) <= tol
    assert nerr(y2, yo) > 3 * tol, nerr(y2, yo)                  # and the result really moved


@pytest.mark.parametrize("dtype,f,tol", [(torch.bfloat16, 512, 2e-2), (torch.float32, 64, 1e-5)])
def test_graphed_forward_only_slice_matches_eager(R, dtype, f, tol):
    """Eval path (dgl/eval.py: batch 1 under torch.no_grad()): gnn_slice_auto replays a captured CUDA graph; results equal the
    eager slice bit for bit, for fresh inputs on every replay, and follow weight updates between replays."""
    from robot_check import make_block
    from oracle import gnn_oracle as O
    blk, _ = make_block(R, O, f, True, True)
    blk.cuda()
    gen = torch.Generator(device="cuda").manual_seed(5)
    for G, K in ((1, 5), (8, 5), (1, 5)):
        g = R.robot_batch(G, K, device="cuda")
        h = torch.rand((G * K, f, 8, 8), generator=gen, device="cuda").to(dtype)
        g.edata["pose"] = torch.randn((G * K * (K - 1), 9), generator=gen, device="cuda")
        with torch.no_grad():
            n0 = R._lib.lib().rgnn_launch_count()
            y = R.gnn_slice_auto(blk, g, h).clone()
            y2 = R.gnn_slice_auto(blk, g, h).clone()             # pure replay: no launches through the C ABI
            n1 = R._lib.lib().rgnn_launch_count()
            ye = R.gnn_slice(blk, g, h)
        assert torch.equal(y, ye) and torch.equal(y2, ye)
    with torch.no_grad():
        for p in blk.parameters():
            p.mul_(0.5)
        y3 = R.gnn_slice_auto(blk, g, h).clone()
        ye3 = R.gnn_slice(blk, g, h)
    assert torch.equal(y3, ye3) and not torch.equal(y3, ye)


@pytest.mark.parametrize("dtype,f,tol", [(torch.bfloat16, 512, 2e-2), (torch.float32, 64, 1e-5)])
def test_graphed_training_slice_inside_an_ordinary_autograd_loop(R, dtype, f, tol):
    """Training at the reference's batch (8 graphs, dgl/training.py:28): gnn_slice_auto with gradients enabled runs a forward and a
    backward CUDA graph behind one autograd node; outputs, dh and every parameter gradient equal the eager path, weight updates
    between steps are picked up, and gradients accumulate into .grad like any autograd node's."""
    from robot_check import make_block
    from oracle import gnn_oracle as O
    import sys
    eng = sys.modules[R.GraphedForward.__module__]
    G, K = 8, 5
    blk, _ = make_block(R, O, f, True, True)
    blk.cuda()
    g = R.robot_batch(G, K, device="cuda")
    gen = torch.Generator(device="cuda").manual_seed(6)
    opt_ = torch.optim.SGD(blk.parameters(), lr=0.05)
    old_flag, eng.GRAPHED_TRAIN = eng.GRAPHED_TRAIN, True
    for it in range(3):
        h = torch.rand((G * K, f, 8, 8), generator=gen, device="cuda").to(dtype)
        pose = torch.randn((G * K * (K - 1), 9), generator=gen, device="cuda")
        dy = torch.randn((G * K, f, 8, 8), generator=gen, device="cuda").to(dtype)
        g.edata["pose"] = pose
        he = h.clone().requires_grad_()
        opt_.zero_grad(set_to_none=True)
        ye = R.gnn_slice(blk, g, he)
        ye.backward(dy)
        ref = {n: p.grad.clone() for n, p in blk.named_parameters()}
        ye = ye.detach()      # drop the eager autograd graph: its AccumulateGrad nodes belong to the default stream and must not be
        #                       alive while the backward graph is captured (torch.cuda.make_graphed_callables' documented rule)
        hg = h.clone().requires_grad_()
        opt_.zero_grad(set_to_none=True)
        n0 = R._lib.lib().rgnn_launch_count()
        yg = R.gnn_slice_auto(blk, g, hg)
        (yg * 1.0).backward(dy)                                  # an op behind the slice, like the decoder
        launched = R._lib.lib().rgnn_launch_count() - n0
        assert it == 0 or launched == 0                          # replays launch nothing through the C ABI
        assert torch.equal(yg, ye) and torch.equal(hg.grad, he.grad)
        for n, p in blk.named_parameters():
            assert nerr(p.grad, ref[n]) <= 1e-5, n               # (atomics in the small gradient kernels reorder fp32 sums)
        opt_.step()                                              # the next replay must see the new weights
    eng.GRAPHED_TRAIN = old_flag
