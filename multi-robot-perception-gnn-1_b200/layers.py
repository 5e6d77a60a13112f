"""Drop-in modules for the graph message-passing layers of the reference's dgl/model/models.py.

Same class names, constructor arguments, forward signatures and state_dict keys as the reference; the
nn.Linear / nn.Conv2d children exist only as parameter containers (default init + checkpoint compatibility) -
the arithmetic runs in librobotgnn's CUDA kernels through functional.py.
"""
import math

import torch
import torch.nn as nn

from . import functional as Fn
from . import _lib


class edge_encoder(nn.Module):
    """dgl/model/models.py:142-155.  state_dict: layers.0.{weight [H,9],bias}, layers.2.{weight [2F,H],bias}."""

    def __init__(self, layers_dim):
        super().__init__()
        self.layers_dim = layers_dim
        self.layers = nn.Sequential(
            nn.Linear(9, self.layers_dim[0]),
            nn.ReLU(),
            nn.Linear(layers_dim[0], layers_dim[1] * 2),
            nn.Sigmoid(),
        )

    def film(self, edge, dtype=torch.float32):
        """z = [E, 2F] with gamma/beta interleaved, exactly the memory the reference's Sigmoid writes."""
        l0, l2 = self.layers[0], self.layers[2]
        return Fn.edge_mlp(edge, l0.weight, l0.bias, l2.weight, l2.bias, dtype)

    def forward(self, edge):
        z = self.film(edge).view(-1, self.layers_dim[1], 2)                     # models.py:154
        return z[:, :, 0].unsqueeze(-1).unsqueeze(-1), z[:, :, 1].unsqueeze(-1).unsqueeze(-1)


class GCN(nn.Module):
    """FiLM-modulated mean aggregation - dgl/model/models.py:213-226 (edge_udf :210-211, node_udf :207-208).

    forward(g) as in the reference (features from g.ndata['image'], poses from g.edata['pose']); also
    forward(g, feat[, pose]) in DGL-nn style.  return_key: 'images' = the aggregate node_udf writes (intended,
    dgl/model/dgl_models.py:189); 'image' = literal reference behaviour, which returns its input (models.py:226).
    film=False gives the update_all(copy_u, mean) variant (models.py:225, dgl_models.py:126).
    """

    def __init__(self, opt, return_key="images", film=True):
        super().__init__()
        if return_key not in ("images", "image"):
            raise ValueError("return_key must be 'images' or 'image'")
        self.opt = opt
        self.return_key, self.film = return_key, film
        self.edge_encoder = edge_encoder(layers_dim=[opt.feature_dim, opt.feature_dim])

    def forward(self, g, feat=None, pose=None):
        with g.local_scope():
            h = g.ndata['image'] if feat is None else feat
            if self.return_key == "image":
                return h
            z = None
            if self.film:
                pose = g.edata['pose'] if pose is None else pose
                z = self.edge_encoder.film(pose, h.dtype)
            return Fn.film_mean(g, h, z)


class GNNBlock(nn.Module):
    """The GNN slice of multi_view_dgl_model (dgl/model/models.py:162-171 ctor, :180-189 forward):
    gcn1 -> cat -> [conv1] -> [gcn2 -> cat -> conv2].  state_dict keys: gcn1.edge_encoder.layers.{0,2}.*,
    conv1.{weight [F,2F,1,1],bias}, and the same under gcn2./conv2. with multi_gcn."""

    def __init__(self, opt, return_key="images", film=True):
        super().__init__()
        self.opt = opt
        f = opt.feature_dim
        self.gcn1 = GCN(opt, return_key, film)
        if getattr(opt, "compress_gcn", False):
            self.conv1 = nn.Conv2d(f * 2, f, kernel_size=1)
        if getattr(opt, "multi_gcn", False):
            assert opt.compress_gcn                                            # models.py:169
            self.gcn2 = GCN(opt, return_key, film)
            self.conv2 = nn.Conv2d(f * 2, f, kernel_size=1)

    def forward(self, g, h, pose=None):
        return gnn_slice(self, g, h, pose)


def _gnn_layer(g, h, pose, gcn, conv):
    if gcn.return_key == "images" and gcn.film:       # the whole layer as one fused autograd node
        l0, l2 = gcn.edge_encoder.layers[0], gcn.edge_encoder.layers[2]
        return Fn.gnn_layer(g, h, pose, l0.weight, l0.bias, l2.weight, l2.bias, conv.weight, conv.bias)
    return Fn.compress(h, gcn(g, h, pose), conv.weight, conv.bias)


def gnn_slice(mod, g, h, pose=None, compute_dtype=None, keep_format=True):
    """dgl/model/models.py:180-189 on any module that owns gcn1[/conv1][/gcn2/conv2] and .opt (GNNBlock, or a
    multi_view_dgl_model that keeps the reference's flat parameter names).

    Seams: `h` may arrive channels-last (a backbone run in torch.channels_last) and / or in another dtype than
    `compute_dtype` (e.g. an fp32 encoder feeding the bf16 throughput mode); it is brought into the kernels' NCHW layout in ONE pass
    (transpose + cast fused), and with keep_format the result goes back to the caller's memory format and dtype the same way, so the
    decoder's first conv (blocks.py:16) gets the layout it was handed before."""
    opt = mod.opt
    pose = g.edata['pose'] if pose is None else pose
    cl_in, dt_in = Fn._is_channels_last(h), h.dtype
    if cl_in or (compute_dtype is not None and compute_dtype != dt_in):
        y = _gnn_slice_nchw(mod, g, Fn.to_gnn_layout(h, compute_dtype or dt_in), pose)
        return Fn.from_gnn_layout(y, dt_in, cl_in) if keep_format else y
    return _gnn_slice_nchw(mod, g, h, pose)


def _gnn_slice_nchw(mod, g, h, pose):
    opt = mod.opt
    if opt.compress_gcn:
        h = _gnn_layer(g, h, pose, mod.gcn1, mod.conv1)
    elif mod.gcn1.return_key == "image":
        h = torch.cat((h, h), dim=1)                                       # literal reference
    else:                                                                   # cat(h, g_h) fused into the aggregation
        z = mod.gcn1.edge_encoder.film(pose, h.dtype) if mod.gcn1.film else None
        h = Fn.film_mean_cat(g, h, z)
    if getattr(opt, "multi_gcn", False):
        h = _gnn_layer(g, h, pose, mod.gcn2, mod.conv2)
    return h


class GraphConv(nn.Module):
    """dgl.nn.pytorch.GraphConv (imported at dgl/model/models.py:7; formula SURVEY App. B).  matmul acts on the
    last feature dim; parameters weight [in,out] (xavier_uniform), bias [out] (zeros)."""

    def __init__(self, in_feats, out_feats, norm='both', weight=True, bias=True, activation=None, allow_zero_in_degree=False):
        super().__init__()
        if norm not in ('none', 'both', 'right', 'left'):
            raise ValueError('Invalid norm value. Must be either "none", "both", "right" or "left". But got "{}".'.format(norm))
        self._in_feats, self._out_feats, self._norm = in_feats, out_feats, norm
        self._allow_zero_in_degree = allow_zero_in_degree
        if weight:
            self.weight = nn.Parameter(torch.empty(in_feats, out_feats))
        else:
            self.register_parameter('weight', None)
        if bias:
            self.bias = nn.Parameter(torch.empty(out_feats))
        else:
            self.register_parameter('bias', None)
        self._activation = activation
        self.reset_parameters()

    def reset_parameters(self):
        if self.weight is not None:
            nn.init.xavier_uniform_(self.weight)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def set_allow_zero_in_degree(self, set_value):
        self._allow_zero_in_degree = set_value

    def forward(self, graph, feat, weight=None, edge_weight=None):
        ix = graph.index()
        if not self._allow_zero_in_degree and _has_zero_in_degree(ix):
            raise RuntimeError('There are 0-in-degree nodes in the graph, output for those nodes will be invalid. '
                               'Adding self-loop on the input graph resolves the issue; or set allow_zero_in_degree=True.')
        if weight is not None and self.weight is not None:
            raise RuntimeError('External weight is provided while at the same time the module has defined its own weight parameter.')
        weight = self.weight if weight is None else weight
        n = ix.N
        inv_out = None
        if self._norm == 'both':
            ns, ds = ix.rsqrt_out, ix.rsqrt_in
        elif self._norm == 'left':
            ns, ds = _inv_out_deg(ix), None
        elif self._norm == 'right':
            ns, ds = None, _inv_in_deg(ix)
        else:
            ns, ds = None, None
        bias_in_gemm = None
        if weight is not None and self._in_feats > self._out_feats:      # multiply first, then aggregate
            feat = Fn.linear(feat, weight)
            rst = Fn.copy_u_reduce(graph, feat, "sum", ns, ds, edge_weight)
            if self.bias is not None:
                rst = Fn.bias_add(rst, self.bias)
        else:
            rst = Fn.copy_u_reduce(graph, feat, "sum", ns, ds, edge_weight)
            if weight is not None:
                rst = Fn.linear(rst, weight, self.bias)
            elif self.bias is not None:
                rst = Fn.bias_add(rst, self.bias)
        if self._activation is not None:
            rst = self._activation(rst)
        return rst


def _has_zero_in_degree(ix):
    """DGL's 0-in-degree check without a per-call device sync: complete K-robot graphs have none by construction; any other
    batch is inspected ONCE (one device->host read when the index is first used) and the answer is cached on the index, so the
    layers stay capture-safe (CUDA graphs, HostPipeline streams)."""
    if getattr(ix, "has_zero_in_deg", None) is None:
        ix.has_zero_in_deg = False if (ix.complete_k >= 2 and ix.N == ix.G * ix.complete_k) else bool((ix.in_deg == 0).any())
    return ix.has_zero_in_deg


def _reciprocal_clamped(deg):
    return 1.0 / deg.clamp(min=1).to(torch.float32)   # [N] index-side vector, not feature math


def _inv_out_deg(ix):
    """1 / max(out_deg, 1), cached on the graph index (the aggregation recognises the index's own degree norms)."""
    if getattr(ix, "inv_out_deg", None) is None:
        ix.inv_out_deg = _reciprocal_clamped(ix.out_deg)
    return ix.inv_out_deg


def _inv_in_deg(ix):
    """1 / max(in_deg, 1): the builder's 1/in_deg vector when no node is isolated (it stores 0 there), else computed."""
    if getattr(ix, "inv_in_deg_clamped", None) is None:
        ix.inv_in_deg_clamped = ix.inv_in_deg if ix.complete_k else _reciprocal_clamped(ix.in_deg)
    return ix.inv_in_deg_clamped


class GCN2Conv(nn.Module):
    """dgl.nn.pytorch.GCN2Conv (GCNII: initial residual + identity mapping; named by BASELINE.json north_star,
    formula SURVEY App. B).  Parameters weight1 [F,F] (weight2 iff project_initial_features=False), bias [F]."""

    def __init__(self, in_feats, layer, alpha=0.1, lambda_=1, project_initial_features=True,
                 allow_zero_in_degree=False, bias=True, activation=None):
        super().__init__()
        self._in_feats = in_feats
        self._project_initial_features = project_initial_features
        self.alpha, self.lambda_ = alpha, lambda_
        self.beta = math.log(lambda_ / layer + 1)
        self._bias, self._activation = bias, activation
        self._allow_zero_in_degree = allow_zero_in_degree
        self.weight1 = nn.Parameter(torch.empty(in_feats, in_feats))
        if self._project_initial_features:
            self.register_parameter("weight2", None)
        else:
            self.weight2 = nn.Parameter(torch.empty(in_feats, in_feats))
        if self._bias:
            self.bias = nn.Parameter(torch.empty(in_feats))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.normal_(self.weight1)
        if not self._project_initial_features:
            nn.init.normal_(self.weight2)
        if self._bias:
            nn.init.zeros_(self.bias)

    def set_allow_zero_in_degree(self, set_value):
        self._allow_zero_in_degree = set_value

    def forward(self, graph, feat, feat_0, edge_weight=None):
        ix = graph.index()
        if not self._allow_zero_in_degree and _has_zero_in_degree(ix):
            raise RuntimeError('There are 0-in-degree nodes in the graph, output for those nodes will be invalid.')
        if edge_weight is None and self._project_initial_features and 2 <= ix.complete_k <= 8 and ix.N == ix.G * ix.complete_k:
            # complete K-robot graph: D^-1/2 A D^-1/2 X is the mean over the other robots; the alpha mix rides on the same pass
            r = Fn.mean_mix(graph, feat, feat_0, 1 - self.alpha, self.alpha)
            rst = Fn.linear(r, self.weight1, self.bias, alpha=self.beta, ident=1 - self.beta)
            return rst if self._activation is None else self._activation(rst)
        agg = Fn.copy_u_reduce(graph, feat, "sum", ix.rsqrt_in, ix.rsqrt_in, edge_weight)   # D^-1/2 A D^-1/2 X (in-degree both sides)
        if self._project_initial_features:
            r = Fn.axpby(agg, feat_0, 1 - self.alpha, self.alpha)
            rst = Fn.linear(r, self.weight1, self.bias, alpha=self.beta, ident=1 - self.beta)
        else:
            a = Fn.axpby(agg, agg, 1 - self.alpha, 0.0)
            b = Fn.axpby(feat_0, feat_0, self.alpha, 0.0)
            ra = Fn.linear(a, self.weight1, None, alpha=self.beta, ident=1 - self.beta)
            rb = Fn.linear(b, self.weight2, self.bias, alpha=self.beta, ident=1 - self.beta)
            rst = Fn.axpby(ra, rb, 1.0, 1.0)
        if self._activation is not None:
            rst = self._activation(rst)
        return rst
