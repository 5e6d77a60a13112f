"""CPU: the oracle restatement (oracle/gnn_oracle.py) against the golden vectors produced by the UNMODIFIED
reference classes (tests/golden/make_golden.py) - this is what pins the oracle."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_cases, nerr
from oracle import gnn_oracle as O


def test_topology_edge_order():
    # dgl/dataloader.py:88-95: eid(i->j) = i*(K-1) + (j if j<i else j-1)
    for k in (2, 3, 5, 8):
        s, d = O.fully_connected_edges(k)
        assert len(s) == k * (k - 1)
        for e, (i, j) in enumerate(zip(s, d)):
            assert i != j and e == i * (k - 1) + (j if j < i else j - 1)


@pytest.mark.parametrize("ci", range(len(load_cases("csr_cases.npz"))))
def test_batch_and_csr_bit_exact(ci):
    c = load_cases("csr_cases.npz")[ci]
    bne = c["batch_num_edges"].tolist()
    cuts = np.cumsum([0] + bne)
    ss = [c["local_src"][cuts[i]:cuts[i + 1]] for i in range(len(bne))]
    dd = [c["local_dst"][cuts[i]:cuts[i + 1]] for i in range(len(bne))]
    b = O.batch_graphs(ss, dd, c["batch_num_nodes"].tolist())
    for k in ("src", "dst", "node_offsets", "edge_offsets", "batch_num_nodes", "batch_num_edges"):
        assert b[k].dtype == np.int64 and np.array_equal(b[k], c[k]), k
    csr = O.build_csr(b["src"], b["dst"], b["node_offsets"][-1])
    for k, v in csr.items():
        assert np.array_equal(v, c[k]) and v.dtype == c[k].dtype, k
    # structural invariants: stable order, ptr consistency
    for v in range(len(csr["in_deg"])):
        seg = csr["in_eid"][csr["in_ptr"][v]:csr["in_ptr"][v + 1]]
        assert np.all(np.diff(seg) > 0) and np.all(b["dst"][seg] == v)
        seg = csr["out_eid"][csr["out_ptr"][v]:csr["out_ptr"][v + 1]]
        assert np.all(np.diff(seg) > 0) and np.all(b["src"][seg] == v)


@pytest.mark.parametrize("ci", range(len(load_cases("layer_cases.npz"))))
def test_layer_matches_reference_bit_exact(ci):
    c = {k: (torch.from_numpy(v) if v.dtype.kind == "f" else v) for k, v in load_cases("layer_cases.npz")[ci].items()}
    h = c["h"].clone().requires_grad_()
    p = {k: c[k].clone().requires_grad_() for k in ("w1", "b1", "w2", "b2")}
    z = O.edge_mlp(c["pose"], p["w1"], p["b1"], p["w2"], p["b2"])
    assert torch.equal(z.detach(), c["z"])                       # same torch ops as nn.Linear/ReLU/Sigmoid
    agg = O.film_mean(h, z, c["src"], c["dst"])
    assert torch.equal(agg.detach(), c["agg"])                   # same order: mul, add, ascending-eid sum, divide
    agg.backward(c["dagg"])
    assert nerr(h.grad, c["dh"]) <= 1e-6
    for k in ("w1", "b1", "w2", "b2"):
        assert nerr(p[k].grad, c["d" + k]) <= 1e-5, k


@pytest.mark.parametrize("ci", range(len(load_cases("block_cases.npz"))))
def test_block_matches_reference(ci):
    c = load_cases("block_cases.npz")[ci]
    seed, f, k, gN, comp, multi = [int(v) for v in c["cfg"]]
    t = lambda a: torch.from_numpy(np.asarray(a))
    layers = []
    for j in range(2 if multi else 1):
        l = {kk: t(c[f"l{j}_{kk}"]).clone().requires_grad_() for kk in ("w1", "b1", "w2", "b2")}
        if comp:
            l["wc"], l["bc"] = t(c[f"l{j}_wc"]).clone().requires_grad_(), t(c[f"l{j}_bc"]).clone().requires_grad_()
        layers.append(l)
    h = t(c["h"]).clone().requires_grad_()
    y = O.gnn_block(h, t(c["pose"]), c["src"], c["dst"], layers, compress_gcn=bool(comp))
    assert nerr(y, c["y"]) <= 2e-6
    y.backward(t(c["dy"]))
    assert nerr(h.grad, c["dh"]) <= 2e-6
    for j, l in enumerate(layers):
        assert nerr(l["w2"].grad, c[f"grad/gcn{j + 1}.edge_encoder.layers.2.weight"]) <= 1e-5
        assert nerr(l["w1"].grad, c[f"grad/gcn{j + 1}.edge_encoder.layers.0.weight"]) <= 1e-5
        if comp:
            assert nerr(l["wc"].grad.reshape(f, 2 * f, 1, 1), c[f"grad/conv{j + 1}.weight"]) <= 1e-5
            assert nerr(l["bc"].grad, c[f"grad/conv{j + 1}.bias"]) <= 1e-5
    # literal reference (returns its input)
    with torch.no_grad():
        y_lit = O.gnn_block(t(c["h"]), t(c["pose"]), c["src"], c["dst"], [{kk: v.detach() for kk, v in l.items()} for l in layers],
                            compress_gcn=bool(comp), return_key="image")
    assert nerr(y_lit, c["y_literal"]) <= 2e-6


def test_script_shape_sample():
    z = np.load(os.path.join(GOLDEN, "script_shape_cases.npz"))
    f = 512
    sb = O.synthetic_batch(8, 5, f, p=64, seed=1)
    assert abs(float(sb["h"].double().sum()) - float(z[f"f{f}_h_sum"])) < 1e-6 * abs(float(z[f"f{f}_h_sum"]))
    assert abs(float(sb["pose"].double().sum()) - float(z[f"f{f}_pose_sum"])) < 1e-6 * max(1.0, abs(float(z[f"f{f}_pose_sum"])))
    p = O.init_layer_params(f, compress_gcn=True, seed=1)
    y = O.gnn_block(sb["h"].reshape(40, f, 8, 8), sb["pose"], sb["src"], sb["dst"], [p], compress_gcn=True)
    got = y.reshape(-1)[torch.from_numpy(z[f"f{f}_idx"])]
    assert float((got - torch.from_numpy(z[f"f{f}_y_s"])).abs().max()) <= 1e-5 * float(z[f"f{f}_y_absmax"])


def test_relative_pose_golden():
    """dgl/utils.py:80-85 self-check values and a 5-robot edge set in reference edge order."""
    z = np.load(os.path.join(GOLDEN, "pose_cases.npz"))
    assert np.array_equal(O.cal_relative_pose(z["p1"], z["p2"]), z["rel12"])
    assert np.array_equal(O.cal_relative_pose(z["p1"], z["p1"]), z["rel11"])
    assert np.allclose(z["rel11"], [0, 0, 0, 1, 0, 0, 0, 1, 0])
    assert np.allclose(z["rel12"][:3], [3.9544525146484375, 3.4390411376953125, 1.8935203552246094e-03])
    assert np.array_equal(O.relative_pose_edges(z["poses"], 5), z["edges"])


def test_graphconv_gcn2_formula_vs_shim_modules():
    """oracle closed forms vs the DGL-semantics shim modules (update_all path)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "dgl_shim"))
    import dgl
    from dgl.nn.pytorch import GraphConv, GCN2Conv
    gen = torch.Generator().manual_seed(0)
    sb = O.synthetic_batch(3, 4, 8, p=1, seed=0)
    g = dgl.graph((sb["src"], sb["dst"]), num_nodes=12)
    x = torch.randn((12, 3, 10), generator=gen, dtype=torch.float64)
    for norm in ("both", "left", "right", "none"):
        for dout in (6, 14):
            m = GraphConv(10, dout, norm=norm).double()
            torch.nn.init.normal_(m.bias)
            assert nerr(O.graphconv(x, sb["src"], sb["dst"], m.weight, m.bias, norm), m(g, x)) < 1e-12
    x2, x0 = torch.randn((12, 10), generator=gen, dtype=torch.float64), torch.randn((12, 10), generator=gen, dtype=torch.float64)
    for project in (True, False):
        m = GCN2Conv(10, layer=3, alpha=0.2, lambda_=0.5, project_initial_features=project).double()
        ref = m(g, x2, x0)
        got = O.gcn2conv(x2, x0, sb["src"], sb["dst"], m.weight1, m.bias, layer=3, alpha=0.2, lambda_=0.5,
                         weight2=None if project else m.weight2)
        assert nerr(got, ref) < 1e-12


@pytest.mark.parametrize("ci", range(len(load_cases("block_cases.npz"))))
def test_ref_port_matches_golden(ci):
    """oracle/ref_port.py (the CPU baseline bench.py times) reproduces the unmodified reference's outputs."""
    from oracle import ref_port
    c = load_cases("block_cases.npz")[ci]
    seed, f, k, gN, comp, multi = [int(v) for v in c["cfg"]]
    t = lambda a: torch.from_numpy(np.asarray(a))
    blk = ref_port.RefBlock(f, bool(comp), bool(multi))
    sd = {}
    for j in range(2 if multi else 1):
        pre = f"gcn{j + 1}.edge_encoder.layers."
        sd.update({pre + "0.weight": t(c[f"l{j}_w1"]), pre + "0.bias": t(c[f"l{j}_b1"]), pre + "2.weight": t(c[f"l{j}_w2"]), pre + "2.bias": t(c[f"l{j}_b2"])})
        if comp:
            sd.update({f"conv{j + 1}.weight": t(c[f"l{j}_wc"]).reshape(f, 2 * f, 1, 1), f"conv{j + 1}.bias": t(c[f"l{j}_bc"])})
    blk.load_state_dict(sd, strict=True)
    g = ref_port.make_graph(c["src"], c["dst"], gN * k, t(c["pose"]))
    h = t(c["h"]).clone().requires_grad_()
    y = blk(g, h)
    assert nerr(y, c["y"]) <= 1e-6
    y.backward(t(c["dy"]))
    assert nerr(h.grad, c["dh"]) <= 1e-6


@pytest.mark.parametrize("ci", range(len(load_cases("block_cases.npz"))))
def test_compiled_reference_matches_golden(ci):
    """oracle/_ref (the reference's own classes, byte-compiled by oracle/build_ref.py) driven by oracle/ref_runner.py - what
    `bench.py --impl reference` times - reproduces the golden outputs made from the reference sources."""
    from oracle import ref_runner
    if not ref_runner.available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    dgl, ref = ref_runner.load()
    c = load_cases("block_cases.npz")[ci]
    seed, f, k, gN, comp, multi = [int(v) for v in c["cfg"]]
    t = lambda a: torch.from_numpy(np.asarray(a))
    model = ref_runner.make_model(ref, f, k, bool(comp), bool(multi))
    assert type(model).__name__ == "multi_view_dgl_model" and type(model.gcn1).__module__ == "model.models"
    sd = {}
    for j in range(2 if multi else 1):
        pre = f"gcn{j + 1}.edge_encoder.layers."
        sd.update({pre + "0.weight": t(c[f"l{j}_w1"]), pre + "0.bias": t(c[f"l{j}_b1"]), pre + "2.weight": t(c[f"l{j}_w2"]), pre + "2.bias": t(c[f"l{j}_b2"])})
        if comp:
            sd.update({f"conv{j + 1}.weight": t(c[f"l{j}_wc"]).reshape(f, 2 * f, 1, 1), f"conv{j + 1}.bias": t(c[f"l{j}_bc"])})
    model.load_state_dict(sd, strict=True)
    g = dgl.graph((t(c["src"]), t(c["dst"])), num_nodes=gN * k)
    g.edata['pose'] = t(c["pose"])
    h = t(c["h"]).clone().requires_grad_()
    y = ref_runner.gnn_slice(ref, model, g, h)
    assert nerr(y, c["y"]) <= 1e-6
    y.backward(t(c["dy"]))
    assert nerr(h.grad, c["dh"]) <= 1e-6
