"""`dgl`-named compatibility package: the DGL surface used by xjh19971/multi-robot-perception-gnn-1, re-hosted on the
B200 robot-graph runtime.  See compat/README.md.  (The reference imports: `from dgl import batch` training.py:15;
`from dgl import load_graphs, graph, save_graphs` dataloader.py:12; `from dgl.data import DGLDataset` dataloader.py:8;
`from dgl.nn.pytorch import GraphConv`, `import dgl.function`, `import dgl.ops` model/models.py:7-9.)"""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

import torch  # noqa: E402

import robot_gnn_b200 as _rg  # noqa: E402

from . import function, ops, data, nn  # noqa: E402,F401

__version__ = "0.6-compat-robot-gnn-b200"

DGLGraph = _rg.RobotGraph
graph = _rg.graph
batch = _rg.batch


def save_graphs(filename, g_list, labels=None):
    """Same call signature as dgl.save_graphs (dataloader.py:169); torch container, not DGL's binary format."""
    g_list = [g_list] if isinstance(g_list, _rg.RobotGraph) else list(g_list)
    payload = [dict(src=g._lsrc.cpu(), dst=g._ldst.cpu(), bnn=g._bnn, bne=g._bne, complete_k=g.complete_k,
                    ndata={k: v.cpu() for k, v in g.ndata.items()}, edata={k: v.cpu() for k, v in g.edata.items()})
               for g in g_list]
    torch.save({"format": "robot-gnn-graphs-v1", "graphs": payload, "labels": labels}, filename)


def load_graphs(filename, idx_list=None):
    """-> (list of graphs, labels dict) like dgl.load_graphs (dataloader.py:174)."""
    try:        # the payload is tensors / lists / dicts / numbers only: the safe unpickler suffices (a tampered cache cannot run code)
        blob = torch.load(filename, weights_only=True)
    except Exception as exc:
        raise ValueError(f"{filename} is not a robot-gnn graph cache (caches written by real DGL must be rebuilt): {exc}") from exc
    if not isinstance(blob, dict) or blob.get("format") != "robot-gnn-graphs-v1":
        raise ValueError(f"{filename} is not a robot-gnn graph cache (caches written by real DGL must be rebuilt)")
    items = blob["graphs"] if idx_list is None else [blob["graphs"][i] for i in idx_list]
    out = []
    for it in items:
        # complete_k is re-derived from the edge lists on the host, never trusted from the file (a wrong value would route the
        # batch to the compiled-topology kernels)
        src, dst, bnn = it["src"].to(torch.int64).reshape(-1), it["dst"].to(torch.int64).reshape(-1), [int(v) for v in it["bnn"]]
        ck = _rg.complete_k_of(src, dst, bnn)
        g = _rg.RobotGraph(src, dst, bnn, it["bne"], ck)
        g.ndata, g.edata = dict(it["ndata"]), dict(it["edata"])
        out.append(g)
    return out, (blob.get("labels") or {})
