"""GPU parity on every BASELINE.json configuration (VERDICT r1 'Next round' item 1): the CUDA path against the fp64 CPU oracle /
the goldens of the unmodified reference at the shapes the configs name - not against itself, and not only through invariants.

  configs[0] C1  train_camera_warehouse_gcn.sh: F=1280, one layer, NO compress, fp32  -> reference golden (script_c1_nocompress.npz)
  configs[1] C2  bench tensors (G=4096, F=512, 2 layers + compress, bf16)             -> first / last 4 graphs vs the oracle
  configs[3] C4  train_camera_warehouse_gcn2_resnet50.sh: F=2048, 2 layers, bf16      -> whole block + all gradients vs the oracle
  configs[4] C5  K in {16, 32, 64} robots (dense complete-graph kernels)              -> whole block + all gradients vs the oracle
  ragged K (a dropped camera in some samples, dgl/dataloader.py:88-99) at F=512 bf16  -> whole block + all gradients vs the oracle
"""
import os
import types

import numpy as np
import pytest
import torch

from conftest import GOLDEN, nerr

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-5
BF16_TOL = 2e-2


def _load(blk, layers, f, comp):
    sd = {}
    for j, l in enumerate(layers):
        pre = f"gcn{j + 1}.edge_encoder.layers."
        sd.update({pre + "0.weight": l["w1"], pre + "0.bias": l["b1"], pre + "2.weight": l["w2"], pre + "2.bias": l["b2"]})
        if comp:
            sd.update({f"conv{j + 1}.weight": l["wc"].reshape(f, 2 * f, 1, 1), f"conv{j + 1}.bias": l["bc"]})
    blk.load_state_dict(sd)
    return blk


def _layers_of(blk, comp=True):
    """The block's own parameters as oracle layer dicts (fp64 copies)."""
    out = []
    for j in (1, 2):
        if not hasattr(blk, f"gcn{j}"):
            break
        enc = getattr(blk, f"gcn{j}").edge_encoder.layers
        d = dict(w1=enc[0].weight, b1=enc[0].bias, w2=enc[2].weight, b2=enc[2].bias)
        if comp:
            conv = getattr(blk, f"conv{j}")
            d.update(wc=conv.weight.reshape(conv.weight.shape[0], -1), bc=conv.bias)
        out.append({k: v.detach().double().cpu() for k, v in d.items()})
    return out


def test_c4_resnet50_width_block_vs_oracle_bf16(R):
    """C4: F=2048, K=5, two layers + compress, bf16 - every tcgen05 load mode (CONV fwd/dX, WGRAD, TN, NN) at the width where the
    gradient all-reduce is 134 MB; output and ALL gradients against the fp64 oracle (train_camera_warehouse_gcn2_resnet50.sh:4)."""
    from robot_check import check_block
    from oracle import gnn_oracle as O
    errs = check_block(R, O, g_=2, k=5, f=2048, dtype=torch.bfloat16, tol=BF16_TOL)
    assert max(errs.values()) <= 2 * BF16_TOL, errs


def test_c1_script_shape_nocompress_vs_reference_golden_fp32(R):
    """C1 exactly as train_camera_warehouse_gcn.sh runs it: mobilenetv2 width F=1280, ONE layer, no compress conv, fp32, G=8, K=5.
    The block output is cat((h, g_h), 1) (models.py:182); golden made by the unmodified reference classes."""
    from oracle import gnn_oracle as O
    z = np.load(os.path.join(GOLDEN, "script_c1_nocompress.npz"))
    f = 1280
    sb = O.synthetic_batch(8, 5, f, p=64, seed=1)
    assert abs(float(sb["h"].double().sum()) - float(z["h_sum"])) < 1e-6 * abs(float(z["h_sum"]))       # generator drift guard
    p = O.init_layer_params(f, compress_gcn=False, seed=1)
    blk = _load(R.GNNBlock(types.SimpleNamespace(feature_dim=f, compress_gcn=False, multi_gcn=False)), [p], f, False).cuda()
    g = R.robot_batch(8, 5, device="cuda")
    g.edata["pose"] = sb["pose"].cuda()
    h = sb["h"].reshape(40, f, 8, 8).cuda().requires_grad_()
    y = blk(g, h)
    assert y.shape == (40, 2 * f, 8, 8)
    dy = torch.randn(y.shape, generator=torch.Generator().manual_seed(101))
    assert abs(float(dy.double().sum()) - float(z["dy_sum"])) < 1e-6 * max(1.0, abs(float(z["dy_sum"])))
    y.backward(dy.cuda())

    def nerr_s(got, ref, idx, scale):
        return float((got.detach().cpu().reshape(-1)[torch.from_numpy(idx)].double() - torch.from_numpy(ref).double()).abs().max() / scale)
    assert nerr_s(y, z["y_s"], z["idx"], float(z["y_absmax"])) <= FP32_TOL
    assert nerr_s(h.grad, z["dh_s"], z["idx_h"], float(z["dh_absmax"])) <= FP32_TOL
    enc = blk.gcn1.edge_encoder.layers
    assert float((enc[2].weight.grad.reshape(-1)[::1009].cpu().double() - torch.from_numpy(z["dw2_s"]).double()).abs().max()) <= 2 * FP32_TOL * float(z["dw2_absmax"])
    assert nerr(enc[0].weight.grad, z["dw1"]) <= 2 * FP32_TOL
    assert nerr(enc[0].bias.grad, z["db1"]) <= 2 * FP32_TOL
    assert nerr(enc[2].bias.grad, z["db2"]) <= 2 * FP32_TOL


@pytest.mark.parametrize("dtype,tol", [(torch.float32, FP32_TOL), (torch.bfloat16, BF16_TOL)])
@pytest.mark.parametrize("g_,k,f", [(3, 16, 128), (2, 32, 128), (1, 64, 128), (2, 64, 64)])
def test_c5_many_robots_block_vs_oracle(R, dtype, tol, g_, k, f):
    """configs[4] sweep widths: K = 16 / 32 / 64 robots per graph run the dense complete-graph kernels (film_dense.cu) and, in bf16,
    the tcgen05 edge-MLP GEMMs over E = G K (K-1) rows; forward and every gradient against the fp64 oracle."""
    from robot_check import check_block
    from oracle import gnn_oracle as O
    errs = check_block(R, O, g_=g_, k=k, f=f, dtype=dtype, tol=tol, multi=False)
    assert max(errs.values()) <= 2 * tol, errs


def test_c2_bench_tensors_first_and_last_graphs_vs_oracle(R):
    """The exact tensors bench.py times (configs[1]: G=4096, K=5, F=512, 2 layers + compress, bf16, seed 1, rank 0): graphs are
    independent, so the block output and dh of the first and the last 4 graphs of THAT run are compared with the fp64 oracle fed
    the same (bf16-rounded) inputs of those graphs and the block's own weights."""
    from oracle import gnn_oracle as O
    G, K, F = 4096, 5, 512
    N, E = G * K, G * K * (K - 1)
    dev = torch.device("cuda")
    torch.manual_seed(1)
    blk = R.GNNBlock(types.SimpleNamespace(feature_dim=F, compress_gcn=True, multi_gcn=True)).to(dev)
    gen = torch.Generator(device=dev).manual_seed(1)
    h = torch.relu(torch.randn((N, F, 8, 8), generator=gen, device=dev, dtype=torch.float32)).bfloat16().requires_grad_()
    dy = torch.randn((N, F, 8, 8), generator=gen, device=dev, dtype=torch.float32).bfloat16()
    pose = torch.randn((E, 9), generator=gen, device=dev, dtype=torch.float32)
    pose[:, :3] *= 3.0
    g = R.robot_batch(G, K, device=dev)
    y = blk(g, h, pose)
    y.backward(dy)
    layers = _layers_of(blk)
    s, d = O.fully_connected_edges(K)
    gs = 4
    b = O.batch_graphs([s] * gs, [d] * gs, [K] * gs)
    for g0 in (0, G - gs):
        n0, n1, e0, e1 = g0 * K, (g0 + gs) * K, g0 * K * (K - 1), (g0 + gs) * K * (K - 1)
        ho = h.detach()[n0:n1].double().cpu().requires_grad_()
        yo = O.gnn_block(ho, pose[e0:e1].double().cpu(), b["src"], b["dst"], layers, compress_gcn=True)
        yo.backward(dy[n0:n1].double().cpu())
        assert nerr(y.detach()[n0:n1].float(), yo) <= BF16_TOL, g0
        assert nerr(h.grad[n0:n1].float(), ho.grad) <= BF16_TOL, g0
    for p in blk.parameters():
        assert p.grad is not None and bool(torch.isfinite(p.grad).all())


@pytest.mark.parametrize("dtype,tol,f", [(torch.bfloat16, BF16_TOL, 512), (torch.float32, FP32_TOL, 64)])
def test_ragged_k_batch_vs_oracle(R, dtype, tol, f):
    """A batch where some samples lost cameras (K in {2..5} mixed; dataloader.py:88-99 builds the graph from the cameras present):
    two layers + compress, output and all gradients against the fp64 oracle."""
    from oracle import gnn_oracle as O
    rng = np.random.RandomState(12)
    ks = [5, 5, 3, 5, 4, 5, 2, 5, 5, 4, 3, 5] + rng.randint(2, 6, size=12).tolist()
    ed = [O.fully_connected_edges(int(k)) for k in ks]
    b = O.batch_graphs([e[0] for e in ed], [e[1] for e in ed], ks)
    n, e = int(b["node_offsets"][-1]), int(b["edge_offsets"][-1])
    gen = torch.Generator().manual_seed(21)
    h32 = torch.relu(torch.randn((n, f, 8, 8), generator=gen))
    pose = torch.randn((e, 9), generator=gen)
    pose[:, :3] *= 3.0
    dy32 = torch.randn((n, f, 8, 8), generator=gen)
    layers = [O.init_layer_params(f, compress_gcn=True, seed=31 + j) for j in range(2)]
    blk = _load(R.GNNBlock(types.SimpleNamespace(feature_dim=f, compress_gcn=True, multi_gcn=True)), layers, f, True).cuda()
    # per-graph dgl.graph + dgl.batch, as the DataLoader collate does (training.py:57-58)
    g = R.batch([R.graph((torch.from_numpy(s_), torch.from_numpy(d_)), num_nodes=int(k)) for (s_, d_), k in zip(ed, ks)]).to("cuda")
    assert g.complete_k == 0 and g.index().buckets is not None and sorted(k for k, _, _ in g.index().buckets) == sorted(set(ks))
    h = h32.cuda().to(dtype).requires_grad_()
    y = blk(g, h, pose.cuda())
    y.backward(dy32.cuda().to(dtype))
    # the per-K bucket launches (compiled-topology ring kernels) and the general CSR-walking kernels agree: bit for bit in fp32
    g2 = R.batch([R.graph((torch.from_numpy(s_), torch.from_numpy(d_)), num_nodes=int(k)) for (s_, d_), k in zip(ed, ks)]).to("cuda")
    g2.index().buckets = None
    h2 = h32.cuda().to(dtype).requires_grad_()
    y2 = blk(g2, h2, pose.cuda())
    y2.backward(dy32.cuda().to(dtype))
    if dtype == torch.float32:
        assert nerr(y2, y) <= 1e-6 and nerr(h2.grad, h.grad) <= 1e-6
    else:
        assert nerr(y2.float(), y.float()) <= tol and nerr(h2.grad.float(), h.grad.float()) <= tol
    for p_ in blk.parameters():
        p_.grad = None
    h.grad = None
    y = blk(g, h, pose.cuda())
    y.backward(dy32.cuda().to(dtype))
    ho = h32.to(dtype).double().requires_grad_()
    ls = [{kk: v.double().requires_grad_() for kk, v in l.items()} for l in layers]
    yo = O.gnn_block(ho, pose.double(), b["src"], b["dst"], ls, compress_gcn=True)
    yo.backward(dy32.to(dtype).double())
    assert nerr(y.float(), yo) <= tol
    assert nerr(h.grad.float(), ho.grad) <= tol
    for j, l in enumerate(ls):
        enc = getattr(blk, f"gcn{j + 1}").edge_encoder.layers
        conv = getattr(blk, f"conv{j + 1}")
        for got, ref, name in ((enc[0].weight.grad, l["w1"].grad, "w1"), (enc[0].bias.grad, l["b1"].grad, "b1"),
                               (enc[2].weight.grad, l["w2"].grad, "w2"), (enc[2].bias.grad, l["b2"].grad, "b2"),
                               (conv.weight.grad.reshape(f, 2 * f), l["wc"].grad, "wc"), (conv.bias.grad, l["bc"].grad, "bc")):
            assert nerr(got, ref) <= 2 * tol, (j, name)
