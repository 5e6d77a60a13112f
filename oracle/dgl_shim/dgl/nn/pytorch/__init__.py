"""dgl.nn.pytorch restated from DGL's published formulas (ORACLE / TEST INFRASTRUCTURE).

``GraphConv`` is imported by the reference (dgl/model/models.py:7) but never
instantiated; ``GCN2Conv`` is named only by BASELINE.json's north_star.  Both are
restated here so the CUDA operators for row A11 have a CPU
checker.  Parity for these two is "unpinned by the reference".
"""
import math

import torch
import torch.nn as nn

from ... import function as fn


class GraphConv(nn.Module):
    def __init__(self, in_feats, out_feats, norm='both', weight=True, bias=True,
                 activation=None, allow_zero_in_degree=False):
        super().__init__()
        if norm not in ('none', 'both', 'right', 'left'):
            raise ValueError('Invalid norm value "{}"'.format(norm))
        self._in_feats, self._out_feats, self._norm = in_feats, out_feats, norm
        self._allow_zero_in_degree = allow_zero_in_degree
        self.weight = nn.Parameter(torch.empty(in_feats, out_feats)) if weight else None
        self.bias = nn.Parameter(torch.empty(out_feats)) if bias else None
        self._activation = activation
        self.reset_parameters()

    def reset_parameters(self):
        if self.weight is not None:
            nn.init.xavier_uniform_(self.weight)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def forward(self, graph, feat, weight=None, edge_weight=None):
        assert edge_weight is None
        with graph.local_scope():
            if not self._allow_zero_in_degree and (graph.in_degrees() == 0).any():
                raise RuntimeError('There are 0-in-degree nodes in the graph')
            if self._norm in ('left', 'both'):
                degs = graph.out_degrees().to(feat).clamp(min=1)
                norm = torch.pow(degs, -0.5) if self._norm == 'both' else 1.0 / degs
                feat = feat * norm.reshape((-1,) + (1,) * (feat.dim() - 1))
            weight = self.weight if weight is None else weight
            if self._in_feats > self._out_feats:
                if weight is not None:
                    feat = torch.matmul(feat, weight)
                graph.ndata['h'] = feat
                graph.update_all(fn.copy_u('h', 'm'), fn.sum('m', 'h'))
                rst = graph.ndata['h']
            else:
                graph.ndata['h'] = feat
                graph.update_all(fn.copy_u('h', 'm'), fn.sum('m', 'h'))
                rst = graph.ndata['h']
                if weight is not None:
                    rst = torch.matmul(rst, weight)
            if self._norm in ('right', 'both'):
                degs = graph.in_degrees().to(feat).clamp(min=1)
                norm = torch.pow(degs, -0.5) if self._norm == 'both' else 1.0 / degs
                rst = rst * norm.reshape((-1,) + (1,) * (feat.dim() - 1))
            if self.bias is not None:
                rst = rst + self.bias
            if self._activation is not None:
                rst = self._activation(rst)
            return rst


class GCN2Conv(nn.Module):
    def __init__(self, in_feats, layer, alpha=0.1, lambda_=1, project_initial_features=True,
                 allow_zero_in_degree=False, bias=True, activation=None):
        super().__init__()
        self._in_feats, self._project = in_feats, project_initial_features
        self.alpha, self.lambda_ = alpha, lambda_
        self.beta = math.log(lambda_ / layer + 1)
        self._allow_zero_in_degree = allow_zero_in_degree
        self._activation = activation
        self.bias = nn.Parameter(torch.empty(in_feats)) if bias else None
        self.weight1 = nn.Parameter(torch.empty(in_feats, in_feats))
        if not project_initial_features:
            self.weight2 = nn.Parameter(torch.empty(in_feats, in_feats))
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.normal_(self.weight1)
        if not self._project:
            nn.init.normal_(self.weight2)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def forward(self, graph, feat, feat_0, edge_weight=None):
        assert edge_weight is None
        with graph.local_scope():
            if not self._allow_zero_in_degree and (graph.in_degrees() == 0).any():
                raise RuntimeError('There are 0-in-degree nodes in the graph')
            degs = graph.in_degrees().to(feat).clamp(min=1)
            norm = torch.pow(degs, -0.5).reshape((-1,) + (1,) * (feat.dim() - 1))
            feat = feat * norm
            graph.ndata['h'] = feat
            graph.update_all(fn.copy_u('h', 'm'), fn.sum('m', 'h'))
            feat = graph.ndata['h'] * norm
            feat = feat * (1 - self.alpha)
            feat_0 = feat_0 * self.alpha
            if self._project:
                rst = feat + feat_0
                rst = (1 - self.beta) * rst + self.beta * torch.matmul(rst, self.weight1)
            else:
                rst = ((1 - self.beta) * feat + self.beta * torch.matmul(feat, self.weight1)
                       + (1 - self.beta) * feat_0 + self.beta * torch.matmul(feat_0, self.weight2))
            if self.bias is not None:
                rst = rst + self.bias
            if self._activation is not None:
                rst = self._activation(rst)
            return rst
