"""DGL-semantics shim (ORACLE / TEST INFRASTRUCTURE - not the product).

A minimal, CPU, pure-PyTorch restatement of the subset of DGL >= 0.5 that
xjh19971/multi-robot-perception-gnn-1 touches, so that the reference's
``dgl/model/models.py`` can be imported and executed UNMODIFIED as an
independent oracle (real DGL cannot be installed in this environment).

Reference call sites restated here:
  dgl.graph((src, dst))            dgl/dataloader.py:99
  dgl.batch(list_of_graphs)        dgl/training.py:57-58, dgl/eval.py:54-55
  g.ndata / g.edata                dgl/model/models.py:175,180,222 ; dgl/dataloader.py:112-122
  g.local_scope()                  dgl/model/models.py:174,220
  g.update_all(edge_udf, node_udf) dgl/model/models.py:223
  g.update_all(fn.copy_u, udf)     dgl/model/models.py:225 (commented), dgl/model/dgl_models.py:126
  g.to(device)                     dgl/training.py:183

Semantics restated (from DGL's published behaviour):
  * idtype int64; num_nodes = max id + 1; edge id = position in the input lists;
    multi-edges allowed.
  * batch(): node ids of graph i shifted by sum_{j<i} num_nodes_j; edges
    concatenated in graph order; every ndata/edata key concatenated on dim 0.
  * update_all with UDFs: messages computed for ALL edges at once; reduction by
    DEGREE BUCKETING - for each distinct in-degree d>0 the mailbox is
    [n_d, d, ...] with each node's messages ordered by ascending edge id;
    zero in-degree nodes receive zeros.
"""
from contextlib import contextmanager

import torch

from . import function  # noqa: F401
from . import ops  # noqa: F401

__version__ = "0.5-semantics-shim"


class _EdgeBatch:
    def __init__(self, g):
        self._g = g
        self.data = g.edata
        self.src = {k: v.index_select(0, g._src) for k, v in g.ndata.items()}
        self.dst = {k: v.index_select(0, g._dst) for k, v in g.ndata.items()}


class _NodeBatch:
    def __init__(self, data, mailbox):
        self.data = data
        self.mailbox = mailbox


class DGLGraph:
    def __init__(self, src, dst, num_nodes=None):
        src = torch.as_tensor(src, dtype=torch.int64).reshape(-1)
        dst = torch.as_tensor(dst, dtype=torch.int64).reshape(-1)
        assert src.shape == dst.shape
        if num_nodes is None:
            num_nodes = int(max(src.max().item(), dst.max().item())) + 1 if src.numel() else 0
        self._src, self._dst, self._n = src, dst, int(num_nodes)
        self.ndata, self.edata = {}, {}
        self._bnn = torch.tensor([self._n], dtype=torch.int64)
        self._bne = torch.tensor([src.numel()], dtype=torch.int64)

    # ---- structure queries -------------------------------------------------
    def num_nodes(self):
        return self._n

    number_of_nodes = num_nodes

    def num_edges(self):
        return int(self._src.numel())

    number_of_edges = num_edges

    def edges(self):
        return self._src, self._dst

    def in_degrees(self):
        return torch.bincount(self._dst, minlength=self._n)

    def out_degrees(self):
        return torch.bincount(self._src, minlength=self._n)

    def batch_num_nodes(self):
        return self._bnn

    def batch_num_edges(self):
        return self._bne

    @property
    def batch_size(self):
        return int(self._bnn.numel())

    def to(self, device):
        g = DGLGraph(self._src.to(device), self._dst.to(device), self._n)
        g.ndata = {k: v.to(device) for k, v in self.ndata.items()}
        g.edata = {k: v.to(device) for k, v in self.edata.items()}
        g._bnn, g._bne = self._bnn, self._bne
        return g

    # ---- frames --------------------------------------------------------------
    @contextmanager
    def local_scope(self):
        old_n, old_e = self.ndata, self.edata
        self.ndata, self.edata = dict(old_n), dict(old_e)
        try:
            yield
        finally:
            self.ndata, self.edata = old_n, old_e

    # ---- message passing -----------------------------------------------------
    def update_all(self, message_func, reduce_func):
        eb = _EdgeBatch(self)
        msgs = message_func(eb)
        deg = self.in_degrees()
        # stable sort by destination => within a destination, ascending edge id
        order = torch.sort(self._dst, stable=True).indices
        indptr = torch.zeros(self._n + 1, dtype=torch.int64)
        indptr[1:] = torch.cumsum(deg, 0)
        out = {}
        for d in torch.unique(deg).tolist():
            if d == 0:
                continue
            nodes = torch.nonzero(deg == d).reshape(-1)
            # [n_d, d] edge ids, ascending within each node
            eids = order[(indptr[nodes].unsqueeze(1) + torch.arange(d).unsqueeze(0))]
            mailbox = {k: v.index_select(0, eids.reshape(-1)).reshape((nodes.numel(), d) + tuple(v.shape[1:]))
                       for k, v in msgs.items()}
            res = reduce_func(_NodeBatch({k: v.index_select(0, nodes) for k, v in self.ndata.items()}, mailbox))
            for k, v in res.items():
                if k not in out:
                    out[k] = torch.zeros((self._n,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
                out[k] = out[k].index_copy(0, nodes, v)
        for k, v in out.items():
            self.ndata[k] = v


def graph(data, num_nodes=None, **_):
    src, dst = data
    return DGLGraph(src, dst, num_nodes)


def batch(graphs):
    noff = eoff = 0
    srcs, dsts = [], []
    for g in graphs:
        srcs.append(g._src + noff)
        dsts.append(g._dst + noff)
        noff += g.num_nodes()
        eoff += g.num_edges()
    bg = DGLGraph(torch.cat(srcs), torch.cat(dsts), noff)
    for k in graphs[0].ndata:
        bg.ndata[k] = torch.cat([g.ndata[k] for g in graphs], 0)
    for k in graphs[0].edata:
        bg.edata[k] = torch.cat([g.edata[k] for g in graphs], 0)
    bg._bnn = torch.cat([g._bnn for g in graphs])
    bg._bne = torch.cat([g._bne for g in graphs])
    return bg
