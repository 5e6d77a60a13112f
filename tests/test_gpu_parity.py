"""GPU parity tests proper: the CUDA path (through the C ABI) vs the golden vectors produced by the UNMODIFIED
reference classes, and vs the CPU oracle on seeded inputs.  Tolerances (BASELINE.json north_star):
bit-exact for CSR / degree / batch indexing; <= 1e-5 normalised error in fp32; <= 2e-2 in bf16."""
import types

import numpy as np
import pytest
import torch

from conftest import load_cases, nerr

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-5
BF16_TOL = 2e-2


def cuda(x, dtype=None):
    t = torch.as_tensor(np.asarray(x)) if not isinstance(x, torch.Tensor) else x
    t = t.cuda()
    return t.to(dtype) if dtype is not None else t


def split_local(case):
    bnn, bne = case["batch_num_nodes"].tolist(), case["batch_num_edges"].tolist()
    return bnn, bne


# ------------------------------------------------------------------------------------------ A1-A3
@pytest.mark.parametrize("ci", range(len(load_cases("csr_cases.npz"))))
def test_csr_bit_exact(R, ci):
    c = load_cases("csr_cases.npz")[ci]
    bnn, bne = split_local(c)
    g = R.RobotGraph(cuda(c["local_src"]), cuda(c["local_dst"]), bnn, bne)
    src, dst = g.edges()
    assert torch.equal(src.cpu(), torch.from_numpy(c["src"])) and torch.equal(dst.cpu(), torch.from_numpy(c["dst"]))
    ix = g.index()
    g.validate()
    assert torch.equal(ix.node_offsets.cpu(), torch.from_numpy(c["node_offsets"]))
    assert torch.equal(ix.edge_offsets.cpu(), torch.from_numpy(c["edge_offsets"]))
    for k in ("in_deg", "out_deg", "in_ptr", "in_src", "in_eid", "out_ptr", "out_dst", "out_eid"):
        got = getattr(ix, k).cpu()
        assert got.dtype == torch.int64
        assert torch.equal(got, torch.from_numpy(c[k])), k
    for k32, k in (("in_ptr32", "in_ptr"), ("out_ptr32", "out_ptr"), ("in_src32", "in_src"), ("in_eid32", "in_eid"),
                   ("out_dst32", "out_dst"), ("out_eid32", "out_eid"), ("src32", "src"), ("dst32", "dst"), ("noff32", "node_offsets")):
        assert torch.equal(getattr(ix, k32).cpu().long(), torch.from_numpy(c[k])), k32
    for k in ("inv_in_deg", "rsqrt_in", "rsqrt_out"):   # fp32, IEEE div/sqrt on both sides -> identical bits
        assert torch.equal(getattr(ix, k).cpu(), torch.from_numpy(c[k])), k


def test_csr_matches_oracle_large(R):
    from oracle import gnn_oracle as O
    rng = np.random.RandomState(5)
    ks = rng.randint(1, 9, size=3000).tolist()
    ed = [O.fully_connected_edges(k) for k in ks]
    b = O.batch_graphs([e[0] for e in ed], [e[1] for e in ed], ks)
    c = O.build_csr(b["src"], b["dst"], b["node_offsets"][-1])
    g = R.RobotGraph(cuda(np.concatenate([e[0] for e in ed])), cuda(np.concatenate([e[1] for e in ed])), ks, [len(e[0]) for e in ed])
    ix = g.index()
    for k in ("in_deg", "out_deg", "in_ptr", "in_src", "in_eid", "out_ptr", "out_dst", "out_eid"):
        assert torch.equal(getattr(ix, k).cpu(), torch.from_numpy(c[k])), k


def test_csr_flags_cross_graph_edge(R):
    g = R.RobotGraph(cuda(np.array([0, 3])), cuda(np.array([1, 0])), [2, 2], [1, 1])   # local id 3 escapes a 2-node graph
    with pytest.raises(ValueError):
        g.validate()


# ------------------------------------------------------------------------------------------ A4-A7, A10
def graph_from_ks(R, ks):
    from oracle import gnn_oracle as O
    ed = [O.fully_connected_edges(int(k)) for k in ks]
    return R.RobotGraph(cuda(np.concatenate([e[0] for e in ed])), cuda(np.concatenate([e[1] for e in ed])),
                        [int(k) for k in ks], [len(e[0]) for e in ed])


@pytest.mark.parametrize("ci", range(len(load_cases("layer_cases.npz"))))
def test_layer_vs_reference_golden_fp32(R, ci):
    c = load_cases("layer_cases.npz")[ci]
    g = graph_from_ks(R, c["ks"])
    f = c["h"].shape[1]
    gcn = R.GCN(types.SimpleNamespace(feature_dim=f)).cuda()
    assert sorted(gcn.state_dict().keys()) == list(c["state_keys"])
    gcn.load_state_dict({"edge_encoder.layers.0.weight": torch.from_numpy(c["w1"]), "edge_encoder.layers.0.bias": torch.from_numpy(c["b1"]),
                         "edge_encoder.layers.2.weight": torch.from_numpy(c["w2"]), "edge_encoder.layers.2.bias": torch.from_numpy(c["b2"])})
    h = cuda(c["h"]).requires_grad_()
    g.ndata["image"], g.edata["pose"] = h, cuda(c["pose"])
    # (1) the FiLM aggregation alone, fed the reference's own gamma/beta: BIT-EXACT (same op order, two roundings, true division)
    agg_exact = R.film_mean(g, h.detach(), cuda(c["z"]))
    assert torch.equal(agg_exact.cpu(), torch.from_numpy(c["agg"]))
    # (2) edge encoder and the whole layer
    gam, bet = gcn.edge_encoder(g.edata["pose"])
    assert gam.shape == (c["pose"].shape[0], f, 1, 1) and gam.stride() == (2 * f, 2, 1, 1)     # strided views, models.py:154-155
    z = torch.stack((gam.reshape(-1, f), bet.reshape(-1, f)), 2).reshape(-1, 2 * f)
    assert nerr(z, c["z"]) <= FP32_TOL
    agg = gcn(g)
    assert nerr(agg, c["agg"]) <= FP32_TOL
    agg.backward(cuda(c["dagg"]))
    assert nerr(h.grad, c["dh"]) <= FP32_TOL
    l0, l2 = gcn.edge_encoder.layers[0], gcn.edge_encoder.layers[2]
    for got, key in ((l0.weight.grad, "dw1"), (l0.bias.grad, "db1"), (l2.weight.grad, "dw2"), (l2.bias.grad, "db2")):
        assert nerr(got, c[key]) <= FP32_TOL, key
    # literal reference behaviour: forward returns its input
    lit = R.GCN(types.SimpleNamespace(feature_dim=f), return_key="image").cuda()
    assert lit(g) is g.ndata["image"]


def load_block(R, c, return_key="images"):
    seed, f, k, gN, comp, multi = [int(v) for v in c["cfg"]]
    opt = types.SimpleNamespace(feature_dim=f, compress_gcn=bool(comp), multi_gcn=bool(multi), camera_num=k)
    blk = R.GNNBlock(opt, return_key=return_key)
    sd = {}
    for j in range(2 if multi else 1):
        pre = f"gcn{j + 1}.edge_encoder.layers."
        sd[pre + "0.weight"], sd[pre + "0.bias"] = torch.from_numpy(c[f"l{j}_w1"]), torch.from_numpy(c[f"l{j}_b1"])
        sd[pre + "2.weight"], sd[pre + "2.bias"] = torch.from_numpy(c[f"l{j}_w2"]), torch.from_numpy(c[f"l{j}_b2"])
        if comp:
            sd[f"conv{j + 1}.weight"] = torch.from_numpy(c[f"l{j}_wc"]).reshape(f, 2 * f, 1, 1)
            sd[f"conv{j + 1}.bias"] = torch.from_numpy(c[f"l{j}_bc"])
    blk.load_state_dict(sd, strict=True)
    return blk.cuda(), (seed, f, k, gN, comp, multi)


@pytest.mark.parametrize("ci", range(len(load_cases("block_cases.npz"))))
def test_block_vs_reference_golden_fp32(R, ci):
    c = load_cases("block_cases.npz")[ci]
    blk, (seed, f, k, gN, comp, multi) = load_block(R, c)
    ref_keys = [s for s in c["state_keys"] if not s.startswith(("encoder", "decoder"))]
    assert sorted(blk.state_dict().keys()) == sorted(ref_keys)
    g = R.robot_batch(gN, k, device="cuda")
    assert torch.equal(g.edges()[0].cpu(), torch.from_numpy(c["src"])) and torch.equal(g.edges()[1].cpu(), torch.from_numpy(c["dst"]))
    g.edata["pose"] = cuda(c["pose"])
    h = cuda(c["h"]).requires_grad_()
    y = blk(g, h)
    assert nerr(y, c["y"]) <= FP32_TOL
    y.backward(cuda(c["dy"]))
    assert nerr(h.grad, c["dh"]) <= FP32_TOL
    for name, p in blk.named_parameters():
        assert nerr(p.grad, c["grad/" + name]) <= 2 * FP32_TOL, name
    # literal mode reproduces the unmodified forward end to end
    lit, _ = load_block(R, c, return_key="image")
    with torch.no_grad():
        assert nerr(lit(g, h.detach()), c["y_literal"]) <= FP32_TOL


@pytest.mark.parametrize("f", [512, 1280])
def test_script_shape_vs_reference_golden_fp32(R, f):
    """G=8, K=5, P=8x8 (dgl/training.py:28, -camera_idx 01234) at resnet18 / mobilenetv2 widths, one layer + compress."""
    from oracle import gnn_oracle as O
    z = np.load(__import__("os").path.join(__import__("conftest").GOLDEN, "script_shape_cases.npz"))
    sb = O.synthetic_batch(8, 5, f, p=64, seed=1)
    p = O.init_layer_params(f, compress_gcn=True, seed=1)
    assert abs(float(sb["h"].double().sum()) - float(z[f"f{f}_h_sum"])) < 1e-6 * abs(float(z[f"f{f}_h_sum"]))   # generator drift guard
    opt = types.SimpleNamespace(feature_dim=f, compress_gcn=True, multi_gcn=False)
    blk = R.GNNBlock(opt)
    blk.load_state_dict({"gcn1.edge_encoder.layers.0.weight": p["w1"], "gcn1.edge_encoder.layers.0.bias": p["b1"],
                         "gcn1.edge_encoder.layers.2.weight": p["w2"], "gcn1.edge_encoder.layers.2.bias": p["b2"],
                         "conv1.weight": p["wc"].reshape(f, 2 * f, 1, 1), "conv1.bias": p["bc"]})
    blk.cuda()
    g = R.robot_batch(8, 5, device="cuda")
    g.edata["pose"] = sb["pose"].cuda()
    h = sb["h"].reshape(40, f, 8, 8).cuda().requires_grad_()
    y = blk(g, h)
    y.backward(sb["dy"].reshape(y.shape).cuda())
    idx = torch.from_numpy(z[f"f{f}_idx"])

    def nerr_s(got, ref, scale):
        return float((got.detach().cpu().reshape(-1)[idx].double() - torch.from_numpy(ref).double()).abs().max() / scale)
    assert nerr_s(y, z[f"f{f}_y_s"], float(z[f"f{f}_y_absmax"])) <= FP32_TOL
    assert nerr_s(h.grad, z[f"f{f}_dh_s"], float(z[f"f{f}_dh_absmax"])) <= FP32_TOL
    assert nerr(blk.conv1.weight.grad.reshape(-1)[::1009], z[f"f{f}_dwc_s"]) <= 2 * FP32_TOL
    assert nerr(blk.gcn1.edge_encoder.layers[2].weight.grad.reshape(-1)[::1009], z[f"f{f}_dw2_s"]) <= 2 * FP32_TOL
    assert nerr(blk.gcn1.edge_encoder.layers[0].weight.grad, z[f"f{f}_dw1"]) <= 2 * FP32_TOL
    assert nerr(blk.gcn1.edge_encoder.layers[2].bias.grad, z[f"f{f}_db2"]) <= 2 * FP32_TOL
    assert nerr(blk.conv1.bias.grad, z[f"f{f}_dbc"]) <= 2 * FP32_TOL


# ------------------------------------------------------------------------------------------ oracle on seeded inputs
def oracle_block(sb, layers, n, f, hw, comp, dtype=torch.float64):
    from oracle import gnn_oracle as O
    h = sb["h"].reshape((n, f) + hw).to(dtype).requires_grad_()
    ls = [{k: v.to(dtype).requires_grad_() for k, v in l.items()} for l in layers]
    y = O.gnn_block(h, sb["pose"].to(dtype), sb["src"], sb["dst"], ls, compress_gcn=comp)
    y.backward(sb["dy_out"].to(dtype))
    return y, h, ls


@pytest.mark.parametrize("dtype,tol", [(torch.float32, FP32_TOL), (torch.bfloat16, BF16_TOL)])
@pytest.mark.parametrize("g_,k,f,comp,multi", [(6, 5, 64, True, True), (4, 8, 128, True, False), (5, 4, 64, False, False), (3, 16, 32, True, True)])
def test_block_vs_oracle(R, dtype, tol, g_, k, f, comp, multi):
    from oracle import gnn_oracle as O
    sb = O.synthetic_batch(g_, k, f, p=64, seed=3)
    n = g_ * k
    layers = [O.init_layer_params(f, compress_gcn=comp, seed=7 + j) for j in range(2 if multi else 1)]
    fo = f if comp else 2 * f
    sb["dy_out"] = torch.randn((n, fo, 8, 8), generator=torch.Generator().manual_seed(11))
    y_o, h_o, ls_o = oracle_block(sb, layers, n, f, (8, 8), comp)
    opt = types.SimpleNamespace(feature_dim=f, compress_gcn=comp, multi_gcn=multi)
    blk = R.GNNBlock(opt)
    sd = {}
    for j, l in enumerate(layers):
        pre = f"gcn{j + 1}.edge_encoder.layers."
        sd.update({pre + "0.weight": l["w1"], pre + "0.bias": l["b1"], pre + "2.weight": l["w2"], pre + "2.bias": l["b2"]})
        if comp:
            sd.update({f"conv{j + 1}.weight": l["wc"].reshape(f, 2 * f, 1, 1), f"conv{j + 1}.bias": l["bc"]})
    blk.load_state_dict(sd)
    blk.cuda()
    g = R.robot_batch(g_, k, device="cuda")
    h = sb["h"].reshape(n, f, 8, 8).cuda().to(dtype).requires_grad_()
    y = blk(g, h, sb["pose"].cuda())
    assert y.dtype == dtype
    y.backward(sb["dy_out"].cuda().to(dtype))
    assert nerr(y.float(), y_o) <= tol
    assert nerr(h.grad.float(), h_o.grad) <= tol
    gtol = 2 * tol
    for j, l in enumerate(ls_o):
        gcn = getattr(blk, f"gcn{j + 1}")
        assert nerr(gcn.edge_encoder.layers[0].weight.grad, l["w1"].grad) <= gtol
        assert nerr(gcn.edge_encoder.layers[0].bias.grad, l["b1"].grad) <= gtol
        assert nerr(gcn.edge_encoder.layers[2].weight.grad, l["w2"].grad) <= gtol
        assert nerr(gcn.edge_encoder.layers[2].bias.grad, l["b2"].grad) <= gtol
        if comp:
            conv = getattr(blk, f"conv{j + 1}")
            assert nerr(conv.weight.grad.reshape(f, 2 * f), l["wc"].grad) <= gtol
            assert nerr(conv.bias.grad, l["bc"].grad) <= gtol


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("k,f", [(5, 48), (3, 64), (8, 32), (12, 40), (20, 24), (33, 16)])
def test_three_aggregation_paths_agree(R, dtype, k, f):
    """complete-graph register kernel, shared-memory staged kernel and global gather kernel: forward bits identical in fp32
    (all keep the reference's rounding order), backward within tolerance; bf16 within 2e-2 of the fp64 oracle."""
    from oracle import gnn_oracle as O
    gN = 7
    sb = O.synthetic_batch(gN, k, f, p=64, seed=9)
    z32 = torch.rand((gN * k * (k - 1), 2 * f), generator=torch.Generator().manual_seed(1))
    dout32 = torch.randn((gN * k, f, 64), generator=torch.Generator().manual_seed(2))
    res = []
    for mode in ("complete", "staged", "generic"):
        g = R.robot_batch(gN, k, device="cuda")
        ix = g.index()
        assert ix.complete_k == k
        if mode != "complete":
            ix.complete_k = 0
        if mode == "generic":
            ix.max_nodes = 0
        h = sb["h"].cuda().to(dtype).requires_grad_()
        z = z32.cuda().to(dtype).requires_grad_()
        a = R.film_mean(g, h, z)
        a.backward(dout32.cuda().to(dtype))
        res.append((a.detach().float(), h.grad.float(), z.grad.float()))
    ho, zo = sb["h"].double().requires_grad_(), z32.double().requires_grad_()
    ref = O.film_mean(ho, zo, sb["src"], sb["dst"])
    ref.backward(dout32.double())
    tol = FP32_TOL if dtype == torch.float32 else BF16_TOL
    for a, dh, dz in res:
        if dtype == torch.float32:
            assert torch.equal(a, res[0][0])
            assert torch.equal(a.cpu(), O.film_mean(sb["h"], z32, sb["src"], sb["dst"]))       # bit-exact vs the fp32 oracle
        assert nerr(a, ref) <= tol and nerr(dh, ho.grad) <= tol and nerr(dz, zo.grad) <= tol


def test_false_completeness_claim_is_caught(R):
    s = torch.tensor([0, 1, 0, 2, 1, 2]).cuda()     # a 3-robot graph whose edges are NOT in the reference enumeration
    d = torch.tensor([1, 0, 2, 0, 2, 1]).cuda()
    g = R.RobotGraph(s, d, [3], [6], complete_k=3)
    with pytest.raises(ValueError, match="complete"):
        g.validate()
    ok = R.robot_batch(4, 3, device="cuda")
    ok.validate()
    host = R.graph(([0, 0, 1, 1, 2, 2], [1, 2, 0, 2, 0, 1]))       # host lists: detected on the host
    assert host.complete_k == 3 and R.graph(([0, 1], [1, 2])).complete_k == 0
    assert R.batch([host, host]).complete_k == 3


def test_copy_u_mean_variant(R):
    """update_all(fn.copy_u('image','m'), mean) - dgl/model/dgl_models.py:126 / models.py:225."""
    from oracle import gnn_oracle as O
    sb = O.synthetic_batch(4, 5, 32, p=64, seed=2)
    g = R.robot_batch(4, 5, device="cuda")
    h = sb["h"].cuda().requires_grad_()
    g.ndata["image"] = h
    g.update_all(R.function.copy_u("image", "m"), R.function.mean("m", "images"))
    out = g.ndata["images"]
    h_o = sb["h"].clone().requires_grad_()
    ref = O.film_mean(h_o, None, sb["src"], sb["dst"])
    assert torch.equal(out.detach().cpu(), ref.detach())
    out.backward(sb["dy"].cuda())
    ref.backward(sb["dy"])
    assert nerr(h.grad, h_o.grad) <= FP32_TOL


# ------------------------------------------------------------------------------------------ A11
@pytest.mark.parametrize("ew", [False, True])
@pytest.mark.parametrize("norm", ["both", "left", "right", "none"])
@pytest.mark.parametrize("din,dout", [(48, 32), (32, 48)])
def test_graphconv_vs_dgl_formula(R, norm, din, dout, ew):
    from oracle import gnn_oracle as O
    gen = torch.Generator().manual_seed(4)
    g_, k = 5, 4
    sb = O.synthetic_batch(g_, k, 8, p=1, seed=4)
    n = g_ * k
    x = torch.randn((n, 6, din), generator=gen)      # [N, P, D]: matmul acts on the last dim
    w = torch.randn((din, dout), generator=gen) * 0.2
    b = torch.randn((dout,), generator=gen)
    dy = torch.randn((n, 6, dout), generator=gen)
    xo, wo, bo = x.double().requires_grad_(), w.double().requires_grad_(), b.double().requires_grad_()
    ewt = torch.rand(len(sb["src"]), generator=gen) + 0.25 if ew else None      # DGL's `edge_weight` argument (u_mul_e messages)
    ref = O.graphconv(xo, sb["src"], sb["dst"], wo, bo, norm, edge_weight=None if ewt is None else ewt.double())
    ref.backward(dy.double())
    conv = R.GraphConv(din, dout, norm=norm)
    assert sorted(conv.state_dict().keys()) == ["bias", "weight"]
    conv.load_state_dict({"weight": w, "bias": b})
    conv.cuda()
    g = R.robot_batch(g_, k, device="cuda")
    xc = x.cuda().requires_grad_()
    y = conv(g, xc, edge_weight=None if ewt is None else ewt.cuda())
    y.backward(dy.cuda())
    assert nerr(y, ref) <= FP32_TOL
    assert nerr(xc.grad, xo.grad) <= FP32_TOL
    assert nerr(conv.weight.grad, wo.grad) <= 2 * FP32_TOL
    assert nerr(conv.bias.grad, bo.grad) <= 2 * FP32_TOL


@pytest.mark.parametrize("ew", [False, True])
@pytest.mark.parametrize("project", [True, False])
def test_gcn2conv_vs_dgl_formula(R, project, ew):
    from oracle import gnn_oracle as O
    gen = torch.Generator().manual_seed(6)
    g_, k, d = 6, 5, 64
    sb = O.synthetic_batch(g_, k, 8, p=1, seed=6)
    n = g_ * k
    x, x0 = torch.randn((n, d), generator=gen), torch.randn((n, d), generator=gen)
    w1, w2 = torch.randn((d, d), generator=gen) * 0.1, torch.randn((d, d), generator=gen) * 0.1
    b = torch.randn((d,), generator=gen)
    dy = torch.randn((n, d), generator=gen)
    t = [v.double().requires_grad_() for v in (x, x0, w1, w2, b)]
    ewt = torch.rand(len(sb["src"]), generator=gen) + 0.25 if ew else None
    ref = O.gcn2conv(t[0], t[1], sb["src"], sb["dst"], t[2], t[4], layer=2, alpha=0.1, lambda_=1.0, weight2=None if project else t[3],
                     edge_weight=None if ewt is None else ewt.double())
    ref.backward(dy.double())
    conv = R.GCN2Conv(d, layer=2, alpha=0.1, lambda_=1, project_initial_features=project)
    sd = {"weight1": w1, "bias": b}
    if not project:
        sd["weight2"] = w2
    assert sorted(conv.state_dict().keys()) == sorted(sd.keys())
    conv.load_state_dict(sd)
    conv.cuda()
    g = R.robot_batch(g_, k, device="cuda")
    xc, x0c = x.cuda().requires_grad_(), x0.cuda().requires_grad_()
    y = conv(g, xc, x0c, edge_weight=None if ewt is None else ewt.cuda())
    y.backward(dy.cuda())
    assert nerr(y, ref) <= FP32_TOL
    assert nerr(xc.grad, t[0].grad) <= FP32_TOL and nerr(x0c.grad, t[1].grad) <= FP32_TOL
    assert nerr(conv.weight1.grad, t[2].grad) <= 2 * FP32_TOL
    assert nerr(conv.bias.grad, t[4].grad) <= 2 * FP32_TOL
    if not project:
        assert nerr(conv.weight2.grad, t[3].grad) <= 2 * FP32_TOL


# ------------------------------------------------------------------------------------------ next-row 3
def test_relative_pose_edges_bit_exact(R):
    """GPU edge-feature producer vs the reference's cal_relative_pose (dgl/utils.py:54-77) golden values, reference edge order."""
    import os
    from conftest import GOLDEN
    from oracle import gnn_oracle as O
    z = np.load(os.path.join(GOLDEN, "pose_cases.npz"))
    got = R.relative_pose_edges(torch.from_numpy(z["poses"]).unsqueeze(0).cuda())
    assert torch.equal(got.cpu(), torch.from_numpy(z["edges"]).float())
    pair = torch.from_numpy(np.stack([z["p1"], z["p2"]])).unsqueeze(0).cuda()           # the reference's own self-check pair
    got2 = R.relative_pose_edges(pair)
    assert torch.equal(got2[0].cpu(), torch.from_numpy(z["rel12"]).float())
    rng = np.random.RandomState(3)                                                        # a batch, several K
    for k in (2, 4, 7):
        poses = rng.randn(6, k, 7)
        poses[..., 3:] /= np.linalg.norm(poses[..., 3:], axis=-1, keepdims=True)
        ref = np.concatenate([O.relative_pose_edges(poses[g], k) for g in range(6)], 0)
        assert torch.equal(R.relative_pose_edges(torch.from_numpy(poses).cuda()).cpu(), torch.from_numpy(ref).float())
