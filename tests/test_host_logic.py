"""CPU: host-side logic of the drop-in boundary - graph bookkeeping, frames, module constructors / state_dict keys,
and the loud failure when asked to compute without CUDA (there is no CPU fallback)."""
import types

import numpy as np
import pytest
import torch

from oracle import gnn_oracle as O


def test_graph_and_batch_bookkeeping(R):
    gs = []
    for k in (3, 5, 2):
        s, d = O.fully_connected_edges(k)
        g = R.graph((s, d))
        assert g.num_nodes() == k and g.num_edges() == k * (k - 1)
        g.ndata["image"] = torch.full((k, 4), float(k))
        g.edata["pose"] = torch.arange(k * (k - 1) * 9, dtype=torch.float32).reshape(-1, 9)
        gs.append(g)
    bg = R.batch(gs)
    assert bg.num_nodes() == 10 and bg.num_edges() == 6 + 20 + 2 and bg.batch_size == 3
    assert bg.batch_num_nodes().tolist() == [3, 5, 2] and bg.batch_num_edges().tolist() == [6, 20, 2]
    ref = O.batch_graphs([O.fully_connected_edges(k)[0] for k in (3, 5, 2)], [O.fully_connected_edges(k)[1] for k in (3, 5, 2)], [3, 5, 2])
    assert np.array_equal(bg.edges()[0].numpy(), ref["src"]) and np.array_equal(bg.edges()[1].numpy(), ref["dst"])
    assert bg.ndata["image"].shape == (10, 4) and bg.ndata["image"][3:8].eq(5).all()
    assert bg.edata["pose"].shape == (28, 9)


def test_robot_batch_matches_reference_enumeration(R):
    g = R.robot_batch(4, 5)
    s, d = O.fully_connected_edges(5)
    b = O.batch_graphs([s] * 4, [d] * 4, [5] * 4)
    assert np.array_equal(g.edges()[0].numpy(), b["src"]) and np.array_equal(g.edges()[1].numpy(), b["dst"])


def test_local_scope_discards_mutations(R):
    g = R.robot_batch(1, 3)
    g.ndata["image"] = torch.zeros(3, 2)
    with g.local_scope():
        g.ndata["image"] = torch.ones(3, 2)
        g.ndata["images"] = torch.ones(3, 2)
        g.edata["pose_gamma"] = torch.ones(6, 1)
        kept = g.ndata["images"]
    assert g.ndata["image"].eq(0).all() and "images" not in g.ndata and "pose_gamma" not in g.edata
    assert kept.eq(1).all()          # tensors created inside stay valid


def test_graph_validation(R):
    with pytest.raises(ValueError):
        R.graph(([0, 1], [1]))
    with pytest.raises(ValueError):
        R.graph(([0, 5], [1, 0]), num_nodes=3)
    with pytest.raises(ValueError):
        R.batch([])
    a, b = R.robot_batch(1, 2), R.robot_batch(1, 2)
    a.ndata["image"] = torch.zeros(2, 1)
    with pytest.raises(ValueError):
        R.batch([a, b])


@pytest.mark.parametrize("comp,multi", [(False, False), (True, False), (True, True)])
def test_state_dict_keys_and_shapes_match_reference(R, comp, multi):
    f = 24
    blk = R.GNNBlock(types.SimpleNamespace(feature_dim=f, compress_gcn=comp, multi_gcn=multi))
    sd = blk.state_dict()
    exp = {"gcn1.edge_encoder.layers.0.weight": (f, 9), "gcn1.edge_encoder.layers.0.bias": (f,),
           "gcn1.edge_encoder.layers.2.weight": (2 * f, f), "gcn1.edge_encoder.layers.2.bias": (2 * f,)}
    if comp:
        exp.update({"conv1.weight": (f, 2 * f, 1, 1), "conv1.bias": (f,)})
    if multi:
        exp.update({k.replace("gcn1", "gcn2").replace("conv1", "conv2"): v for k, v in exp.items()})
    assert {k: tuple(v.shape) for k, v in sd.items()} == exp          # SURVEY §8b, verified against the reference class
    gcn = R.GCN(types.SimpleNamespace(feature_dim=f))
    assert sorted(gcn.state_dict()) == ["edge_encoder.layers.0.bias", "edge_encoder.layers.0.weight",
                                        "edge_encoder.layers.2.bias", "edge_encoder.layers.2.weight"]
    with pytest.raises(AssertionError):                                # models.py:169
        R.GNNBlock(types.SimpleNamespace(feature_dim=f, compress_gcn=False, multi_gcn=True))


def test_dgl_style_module_params(R):
    assert {k: tuple(v.shape) for k, v in R.GraphConv(8, 4).state_dict().items()} == {"weight": (8, 4), "bias": (4,)}
    assert set(R.GCN2Conv(8, 1).state_dict()) == {"weight1", "bias"}
    assert set(R.GCN2Conv(8, 1, project_initial_features=False).state_dict()) == {"weight1", "weight2", "bias"}
    with pytest.raises(ValueError):
        R.GraphConv(8, 4, norm="bogus")


def test_no_cpu_fallback(R):
    """Compute on CPU tensors must fail loudly, never fall back."""
    g = R.robot_batch(2, 3)
    gcn = R.GCN(types.SimpleNamespace(feature_dim=8))
    g.ndata["image"] = torch.zeros(6, 8, 2, 2)
    g.edata["pose"] = torch.zeros(12, 9)
    with pytest.raises(RuntimeError, match="CUDA"):
        gcn(g)
    with pytest.raises(RuntimeError, match="CUDA"):
        g.index()
    lit = R.GCN(types.SimpleNamespace(feature_dim=8), return_key="image")
    assert lit(g) is g.ndata["image"]                                  # literal reference mode does no compute


def test_missing_library_is_loud(R, monkeypatch):
    monkeypatch.setattr(R._lib, "_lib", None)
    monkeypatch.setattr(R._lib, "LIB_PATH", "/nonexistent/librobotgnn.so")
    with pytest.raises(RuntimeError, match="librobotgnn.so not found"):
        R._lib.lib()
