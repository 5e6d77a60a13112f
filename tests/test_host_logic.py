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


def test_per_graph_complete_k_survives_batching_and_drives_the_k_buckets():
    """A ragged batch (samples with different camera counts, dataloader.py:88-99): batch() keeps every graph's host-verified
    complete-K claim, the uniform claim is dropped, and the bucket table groups graphs by K with their node / edge bases."""
    import robot_gnn_b200 as R
    import torch
    ks = [5, 3, 5, 4, 3]
    gs = [R.robot_batch(1, k) for k in ks]
    b = R.batch(gs)
    assert b.complete_k == 0 and b.complete_ks == ks
    assert R.batch([R.robot_batch(1, 4), R.robot_batch(2, 4)]).complete_k == 4
    # a graph whose edge list is not the reference enumeration poisons only itself
    odd = R.graph((torch.tensor([0, 1, 2]), torch.tensor([1, 2, 0])), num_nodes=3)
    assert odd.complete_k == 0 and R.batch([gs[0], odd]).complete_ks == [5, 0]
    buckets = b._k_buckets(torch.device("cpu"))
    assert [k for k, _, _ in buckets] == [3, 4, 5] and [n for _, n, _ in buckets] == [2, 1, 2]
    nb = [0, 5, 8, 13, 17]
    eb = [0, 20, 26, 46, 58]
    want = {3: [nb[1], eb[1], nb[4], eb[4]], 4: [nb[3], eb[3]], 5: [nb[0], eb[0], nb[2], eb[2]]}
    for k, n, t in buckets:
        assert t.dtype == torch.int32 and t.tolist() == want[k]
    assert R.batch([gs[0], odd])._k_buckets(torch.device("cpu")) is None
    assert R.complete_k_of(b._lsrc[:20], b._ldst[:20], [5]) == 5 and R.complete_k_of(b._lsrc, b._ldst, ks) == 0


def test_fused_backward_weight_row_permutation():
    """rgnn_tc_conv_dx_film_bwd wants Wc^T with, per 64-channel tile and group of 8 channels, the 8 direct rows followed by the 8
    aggregated rows of the same channels (include/robot_gnn.h)."""
    import robot_gnn_b200 as R
    import torch
    f = 192
    perm = R.functional._dx_film_perm(f, torch.device("cpu"))
    assert sorted(perm.tolist()) == list(range(2 * f))
    for t in range(f // 64):
        for j in range(8):
            rows = perm[128 * t + 16 * j: 128 * t + 16 * j + 16].tolist()
            assert rows[:8] == [64 * t + 8 * j + i for i in range(8)] and rows[8:] == [f + 64 * t + 8 * j + i for i in range(8)]
