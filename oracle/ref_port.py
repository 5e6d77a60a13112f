"""CPU port of the reference GNN block, op for op, for timing the CPU baseline on the GPU box
(ORACLE / TEST INFRASTRUCTURE; /root/reference does not travel, so bench.py times this port).

It restates, over oracle/dgl_shim, exactly what the reference executes per step:
  edge_encoder  (nn.Linear(9,F) ReLU nn.Linear(F,2F) Sigmoid, view(-1,F,2) split)      dgl/model/models.py:142-155
  GCN.forward   (local_scope, edata gamma/beta, update_all(edge_udf, node_udf))         dgl/model/models.py:207-226
  block         (cat((h, g_h), 1), Conv2d(2F,F,1), optional second layer)              dgl/model/models.py:180-189
with the intended return key 'images' (SURVEY F5).  Validated against the unmodified reference through the
golden vectors (tests/test_oracle_golden.py::test_ref_port_matches_golden).
"""
import os
import sys

import torch
import torch.nn as nn

_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dgl_shim")


def _dgl():
    if _SHIM not in sys.path:
        sys.path.insert(0, _SHIM)
    import dgl
    return dgl


def _node_udf(nodes):
    return {'images': nodes.mailbox['m'].mean(1)}


def _edge_udf(edges):
    return {'m': edges.data['pose_gamma'] * edges.src['image'] + edges.data['pose_beta']}


class RefEdgeEncoder(nn.Module):
    def __init__(self, f):
        super().__init__()
        self.f = f
        self.layers = nn.Sequential(nn.Linear(9, f), nn.ReLU(), nn.Linear(f, 2 * f), nn.Sigmoid())

    def forward(self, edge):
        edge = self.layers(edge.float()).view(-1, self.f, 2)
        return edge[:, :, 0].unsqueeze(-1).unsqueeze(-1), edge[:, :, 1].unsqueeze(-1).unsqueeze(-1)


class RefGCN(nn.Module):
    def __init__(self, f):
        super().__init__()
        self.edge_encoder = RefEdgeEncoder(f)

    def forward(self, g):
        with g.local_scope():
            g.edata['pose_gamma'], g.edata['pose_beta'] = self.edge_encoder(g.edata['pose'])
            g.update_all(_edge_udf, _node_udf)
            return g.ndata['images']


class RefBlock(nn.Module):
    def __init__(self, f, compress_gcn=True, multi_gcn=True):
        super().__init__()
        self.compress_gcn, self.multi_gcn = compress_gcn, multi_gcn
        self.gcn1 = RefGCN(f)
        if compress_gcn:
            self.conv1 = nn.Conv2d(2 * f, f, kernel_size=1)
        if multi_gcn:
            self.gcn2 = RefGCN(f)
            self.conv2 = nn.Conv2d(2 * f, f, kernel_size=1)

    def forward(self, g, h):
        with g.local_scope():
            g.ndata['image'] = h
            g_h = self.gcn1(g)
            h = torch.cat((h, g_h), dim=1)
            if self.compress_gcn:
                h = self.conv1(h)
            if self.multi_gcn:
                g.ndata['image'] = h
                g_h = self.gcn2(g)
                h = torch.cat((h, g_h), dim=1)
                h = self.conv2(h)
            return h


def make_graph(src, dst, num_nodes, pose):
    g = _dgl().graph((torch.as_tensor(src), torch.as_tensor(dst)), num_nodes=num_nodes)
    g.edata['pose'] = pose
    return g


def time_block(f, k, graphs, compress_gcn, multi_gcn, reps=3, threads=None, seed=1):
    """fwd+bwd seconds (best of `reps` after one warm-up) of the port on `graphs` K-robot graphs, fp32, all host threads."""
    import time
    from . import gnn_oracle as O
    if threads:
        torch.set_num_threads(threads)
    sb = O.synthetic_batch(graphs, k, f, p=64, seed=seed)
    n = graphs * k
    blk = RefBlock(f, compress_gcn, multi_gcn)
    g = make_graph(sb["src"], sb["dst"], n, sb["pose"])
    h = sb["h"].reshape(n, f, 8, 8).clone().requires_grad_()
    dy = torch.randn((n, f if compress_gcn else 2 * f, 8, 8), generator=torch.Generator().manual_seed(2))
    best = float("inf")
    for it in range(reps + 1):
        blk.zero_grad(set_to_none=True)
        h.grad = None
        t0 = time.perf_counter()
        y = blk(g, h)
        y.backward(dy)
        dt = time.perf_counter() - t0
        if it > 0:
            best = min(best, dt)
    return best


def time_gcnii(f, k, graphs, layers=2, reps=3, threads=None, seed=1):
    """fwd+bwd seconds (best of `reps` after one warm-up) of a stack of DGL-formula GCN2Conv layers (oracle/dgl_shim, SURVEY
    App. B) on `graphs` complete K-robot graphs, channels-last features [N, 64, F], fp32, all host threads.  The reference imports
    dgl.nn.pytorch but never runs GCN2Conv; this is the CPU counterpart of the north-star GCNII stack."""
    import time
    from . import gnn_oracle as O
    dgl = _dgl()
    from dgl.nn.pytorch import GCN2Conv
    if threads:
        torch.set_num_threads(threads)
    sb = O.synthetic_batch(graphs, k, 8, p=64, seed=seed)        # topology only
    n = graphs * k
    g = dgl.graph((torch.as_tensor(sb["src"]), torch.as_tensor(sb["dst"])), num_nodes=n)
    gen = torch.Generator().manual_seed(seed)
    convs = [GCN2Conv(f, layer=i + 1, alpha=0.1, lambda_=1) for i in range(layers)]
    x0 = torch.randn((n, 64, f), generator=gen).requires_grad_()
    dy = torch.randn((n, 64, f), generator=gen)
    best = float("inf")
    for it in range(reps + 1):
        for c in convs:
            c.zero_grad(set_to_none=True)
        x0.grad = None
        t0 = time.perf_counter()
        x = x0
        for c in convs:
            x = c(g, x, x0)
        x.backward(dy)
        dt = time.perf_counter() - t0
        if it > 0:
            best = min(best, dt)
    return best
