"""Script-level drop-in: the `dgl`-named compatibility package and the `model.models` replacement.
CPU tests cover imports, dataset protocol, graph cache round trip and checkpoint-key compatibility with the reference;
the GPU test runs a miniature of the reference's train_dgl loop (dgl/training.py:176-218)."""
import json
import os
import sys
import types

import pytest
import torch

from conftest import GOLDEN, ROOT


@pytest.fixture()
def compat(monkeypatch):
    """Import compat/dgl as `dgl` and compat/model as `model` (fresh each time; never leaks into other tests)."""
    for k in [k for k in sys.modules if k == "dgl" or k.startswith("dgl.") or k == "model" or k.startswith("model.")]:
        monkeypatch.delitem(sys.modules, k)
    monkeypatch.syspath_prepend(os.path.join(ROOT, "compat"))
    import dgl
    from model import models
    yield dgl, models
    for k in [k for k in sys.modules if k == "dgl" or k.startswith("dgl.") or k == "model" or k.startswith("model.")]:
        sys.modules.pop(k, None)


def make_opt(**kw):
    base = dict(pretrained=False, camera_num=5, image_size=64, task="depth", output_dim=1, backbone="resnet18", feature_dim=512,
                compress_gcn=True, multi_gcn=True, skip_level=False)
    base.update(kw)
    return types.SimpleNamespace(**base)


def test_reference_import_surface(compat):
    dgl, models = compat
    from dgl import batch, graph, load_graphs, save_graphs          # training.py:15, dataloader.py:12
    from dgl.data import DGLDataset                                  # dataloader.py:8
    from dgl.nn.pytorch import GraphConv                              # models.py:7
    import dgl.function as fn                                         # models.py:8
    import dgl.ops                                                    # models.py:9
    assert callable(fn.copy_u) and callable(batch) and callable(graph) and GraphConv is not None and DGLDataset is not None
    for name in ("multi_view_dgl_model", "GCN", "edge_encoder", "encoder", "decoder", "single_view_model"):
        assert hasattr(models, name)


@pytest.mark.parametrize("name,kw", [("resnet18_compress_multi", {}),
                                     ("mobilenetv2_plain_skip", dict(backbone="mobilenetv2", feature_dim=1280, compress_gcn=False,
                                                                     multi_gcn=False, skip_level=True))])
def test_checkpoint_keys_match_reference(compat, name, kw):
    _, models = compat
    golden = json.load(open(os.path.join(GOLDEN, "model_state_keys.json")))[name]
    sd = models.multi_view_dgl_model(make_opt(**kw)).state_dict()
    assert {k: list(v.shape) for k, v in sd.items()} == golden


def test_dataset_protocol_and_graph_cache(compat, tmp_path):
    dgl, _ = compat

    class DS(dgl.data.DGLDataset):
        def __init__(self, save_dir):
            self.calls = []
            super().__init__(name="robots", save_dir=str(save_dir))

        def process(self):
            self.calls.append("process")
            self.graphs = []
            for i in range(3):
                g = dgl.graph(([0, 0, 1, 1, 2, 2], [1, 2, 0, 2, 0, 1]))
                g.ndata["image"] = torch.full((3, 2), float(i))
                g.edata["pose"] = torch.arange(54, dtype=torch.float32).reshape(6, 9) + i
                self.graphs.append(g)

        def save(self):
            self.calls.append("save")
            dgl.save_graphs(os.path.join(self.save_dir, "dgl_graph_3.bin"), self.graphs)

        def load(self):
            self.calls.append("load")
            self.graphs, _ = dgl.load_graphs(os.path.join(self.save_dir, "dgl_graph_3.bin"))

        def has_cache(self):
            return os.path.exists(os.path.join(self.save_dir, "dgl_graph_3.bin"))

        def __getitem__(self, i):
            return self.graphs[i]

        def __len__(self):
            return len(self.graphs)

    a = DS(tmp_path)
    assert a.calls == ["process", "save"]
    b = DS(tmp_path)
    assert b.calls == ["load"] and len(b) == 3
    for ga, gb in zip(a.graphs, b.graphs):
        assert torch.equal(ga.edges()[0], gb.edges()[0]) and torch.equal(ga.ndata["image"], gb.ndata["image"])
        assert torch.equal(ga.edata["pose"], gb.edata["pose"]) and gb.complete_k == 3
    bg = dgl.batch([b[0], b[2]])
    assert bg.num_nodes() == 6 and bg.batch_num_edges().tolist() == [6, 6] and bg.complete_k == 3
    with pytest.raises(ValueError):
        torch.save({"x": 1}, tmp_path / "foreign.bin")
        dgl.load_graphs(str(tmp_path / "foreign.bin"))


@pytest.mark.gpu
def test_miniature_train_loop(compat):
    """dgl/training.py:176-218 in miniature: collate with dgl.batch, data.to('cuda:0'), model(data), loss, backward, Adam step."""
    dgl, models = compat
    torch.manual_seed(0)
    opt = make_opt(image_size=64)
    model = models.multi_view_dgl_model(opt).cuda()
    optim = torch.optim.Adam(model.parameters(), lr=0.005)            # training.py:317
    from oracle import gnn_oracle as O
    s, d = O.fully_connected_edges(5)
    samples = []
    for i in range(4):
        g = dgl.graph((s, d))
        g.ndata["image"] = torch.randn(5, 3, 64, 64)
        g.ndata["depth"] = torch.rand(5, 1, 64, 64)
        g.edata["pose"] = torch.randn(20, 9)
        samples.append(g)
    before = {k: v.detach().clone() for k, v in model.named_parameters() if k.startswith(("gcn", "conv"))}
    losses = []
    for it in range(2):
        data = dgl.batch(samples[2 * it: 2 * it + 2]).to('cuda:0')     # training.py:58,183
        optim.zero_grad()
        depths = data.ndata['depth'].view(-1, opt.camera_num, 1, opt.image_size, opt.image_size)
        pred = model(data)
        assert pred.shape == depths.shape
        loss = torch.nn.functional.smooth_l1_loss(pred, depths)
        loss.backward()
        optim.step()
        losses.append(loss.item())
    assert all(torch.isfinite(torch.tensor(losses)))
    changed = [k for k, v in model.named_parameters() if k in before and not torch.equal(v.detach(), before[k])]
    assert len(changed) == len(before)          # every GNN-layer parameter received a gradient (no dead edge encoder: SURVEY F5)
