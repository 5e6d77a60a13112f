"""Closed-form CPU restatement of the robot-graph message-passing hot path.

ORACLE / TEST INFRASTRUCTURE - never imported by the product package.

Every function cites the reference lines (relative to /root/reference) it
follows.  Integer work is numpy int64 (bit-exact target); floating point work is
torch on CPU in a caller-chosen dtype (float64 for "truth", float32 to mimic
the reference's arithmetic), gradients by autograd.

Pinned by tests/golden/*.npz, which were produced by running the UNMODIFIED
reference classes over oracle/dgl_shim (tests/golden/make_golden.py).
"""
from __future__ import annotations

import math

import numpy as np
import torch

# --------------------------------------------------------------------------------------
# A1  per-sample topology                                    dgl/dataloader.py:88-95,99
# --------------------------------------------------------------------------------------

def fully_connected_edges(k: int):
    """src-major enumeration of the directed K-clique without self loops.

    Follows dgl/dataloader.py:88-95 (``for i: for j: if i != j: x.append(i); y.append(j)``).
    Edge id == list position, i.e. eid(i->j) = i*(K-1) + (j if j < i else j-1).
    """
    src, dst = [], []
    for i in range(k):
        for j in range(k):
            if i != j:
                src.append(i)
                dst.append(j)
    return np.asarray(src, dtype=np.int64), np.asarray(dst, dtype=np.int64)


# --------------------------------------------------------------------------------------
# A2  dgl.batch                                              dgl/training.py:57-58
# --------------------------------------------------------------------------------------

def batch_graphs(srcs, dsts, num_nodes):
    """Restates ``dgl.batch`` (call site dgl/training.py:58) on COO lists.

    srcs/dsts: list of int64 arrays with per-graph LOCAL node ids;
    num_nodes: list of ints.  Returns dict of int64 arrays:
      src, dst (global ids, concatenated in graph order), batch_num_nodes,
      batch_num_edges, node_offsets [G+1], edge_offsets [G+1].
    """
    bnn = np.asarray(num_nodes, dtype=np.int64)
    bne = np.asarray([len(s) for s in srcs], dtype=np.int64)
    noff = np.zeros(len(bnn) + 1, dtype=np.int64)
    eoff = np.zeros(len(bnn) + 1, dtype=np.int64)
    np.cumsum(bnn, out=noff[1:])
    np.cumsum(bne, out=eoff[1:])
    src = np.concatenate([np.asarray(s, dtype=np.int64) + noff[i] for i, s in enumerate(srcs)]) if len(srcs) else np.zeros(0, np.int64)
    dst = np.concatenate([np.asarray(d, dtype=np.int64) + noff[i] for i, d in enumerate(dsts)]) if len(dsts) else np.zeros(0, np.int64)
    return dict(src=src, dst=dst, batch_num_nodes=bnn, batch_num_edges=bne,
                node_offsets=noff, edge_offsets=eoff)


# --------------------------------------------------------------------------------------
# A3  in-degree / in-edge lookup used by update_all          call site dgl/model/models.py:223
# --------------------------------------------------------------------------------------

def build_csr(src, dst, num_nodes):
    """In-CSR ("CSC") and out-CSR of a COO graph, DGL conventions.

    Stable in COO order: inside one destination (source) the in- (out-) edges
    ascend by edge id - that is the mailbox order of DGL's degree bucketing.
    Also returns the normalisation vectors GraphConv / GCN2Conv / the FiLM-mean
    use (fp32): inv_in_deg = 1/in_deg (0 for isolated nodes), rsqrt_in =
    clamp(in_deg,1)^-1/2, rsqrt_out = clamp(out_deg,1)^-1/2.
    """
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    n = int(num_nodes)
    in_deg = np.bincount(dst, minlength=n).astype(np.int64)
    out_deg = np.bincount(src, minlength=n).astype(np.int64)
    in_ptr = np.zeros(n + 1, np.int64)
    out_ptr = np.zeros(n + 1, np.int64)
    np.cumsum(in_deg, out=in_ptr[1:])
    np.cumsum(out_deg, out=out_ptr[1:])
    in_eid = np.argsort(dst, kind="stable").astype(np.int64)
    out_eid = np.argsort(src, kind="stable").astype(np.int64)
    f32 = np.float32
    with np.errstate(divide="ignore"):
        inv_in = np.where(in_deg > 0, f32(1.0) / in_deg.astype(f32), f32(0)).astype(f32)
    rsqrt_in = (f32(1.0) / np.sqrt(np.maximum(in_deg, 1).astype(f32))).astype(f32)
    rsqrt_out = (f32(1.0) / np.sqrt(np.maximum(out_deg, 1).astype(f32))).astype(f32)
    return dict(in_deg=in_deg, out_deg=out_deg,
                in_ptr=in_ptr, in_src=src[in_eid], in_eid=in_eid,
                out_ptr=out_ptr, out_dst=dst[out_eid], out_eid=out_eid,
                inv_in_deg=inv_in, rsqrt_in=rsqrt_in, rsqrt_out=rsqrt_out)


# --------------------------------------------------------------------------------------
# A4  edge encoder                                           dgl/model/models.py:142-155
# --------------------------------------------------------------------------------------

def edge_mlp(pose, w1, b1, w2, b2):
    """Z = sigmoid(relu(pose W1^T + b1) W2^T + b2)  -> [E, 2F]; gamma = Z[:,0::2], beta = Z[:,1::2]
    (the interleave comes from ``view(-1, F, 2)``, dgl/model/models.py:154-155)."""
    a1 = torch.relu(pose @ w1.t() + b1)
    return torch.sigmoid(a1 @ w2.t() + b2)


# --------------------------------------------------------------------------------------
# A5/A6  FiLM message + mean reduce                          dgl/model/models.py:207-211
# --------------------------------------------------------------------------------------

def film_mean(h, z, src, dst, num_nodes=None):
    """agg[v] = (1/deg v) sum_{e: dst_e = v, ascending e} (gamma_e * h[src_e] + beta_e); 0 if deg v = 0.

    h: [N, F, *spatial]; z: [E, 2F] interleaved gamma/beta (or None => gamma=1, beta=0, the
    copy_u-mean variant of dgl/model/dgl_models.py:126).  Two roundings (mul, then add) and an
    ascending-eid sum followed by a division, as the reference does (models.py:211, :208).
    """
    n = h.shape[0] if num_nodes is None else num_nodes
    src = torch.as_tensor(src, dtype=torch.int64)
    dst = torch.as_tensor(dst, dtype=torch.int64)
    f = h.shape[1]
    hs = h.index_select(0, src)
    if z is not None:
        bshape = (z.shape[0], f) + (1,) * (h.dim() - 2)
        m = z[:, 0::2].reshape(bshape) * hs + z[:, 1::2].reshape(bshape)
    else:
        m = hs
    agg = torch.zeros((n,) + tuple(h.shape[1:]), dtype=h.dtype)
    agg = agg.index_add(0, dst, m)  # index_add on CPU accumulates in ascending index order
    deg = torch.bincount(dst, minlength=n).to(h.dtype).clamp(min=1)   # isolated nodes keep their zeros
    return agg / deg.reshape((-1,) + (1,) * (h.dim() - 1))               # ``mean`` = sum / count (a true division)


def compress(h, g_h, wc, bc):
    """Y = conv1x1(cat(h, g_h))   dgl/model/models.py:182-184 (Conv2d(2F, F, 1), models.py:164)."""
    x = torch.cat((h, g_h), 1)
    w = wc.reshape(wc.shape[0], wc.shape[1])
    y = torch.einsum("ok,nk...->no...", w, x)
    return y + bc.reshape((1, -1) + (1,) * (h.dim() - 2))


def gnn_block(h, pose, src, dst, layers, compress_gcn=True, return_key="images", film=True):
    """The GNN slice of multi_view_dgl_model.forward (dgl/model/models.py:180-189).

    layers: list (1 or 2 entries, the 2nd = ``-multi_gcn``) of dicts with keys
    w1,b1,w2,b2 (edge encoder) and, if compress_gcn, wc,bc (conv1/conv2).
    return_key 'images' = intended output of GCN.forward (dgl_models.py:189);
    'image' = the literal reference (models.py:226 returns the input).
    """
    for lyr in layers:
        z = edge_mlp(pose.to(h.dtype), lyr["w1"], lyr["b1"], lyr["w2"], lyr["b2"]) if film else None
        g_h = film_mean(h, z, src, dst) if return_key == "images" else h
        if compress_gcn:
            h = compress(h, g_h, lyr["wc"], lyr["bc"])
        else:
            h = torch.cat((h, g_h), 1)
    return h


# --------------------------------------------------------------------------------------
# A11  north-star operators not executed by the reference (DGL formulas)
# --------------------------------------------------------------------------------------

def graphconv(feat, src, dst, weight, bias, norm="both", edge_weight=None):
    """dgl.nn.pytorch.GraphConv.forward (imported at dgl/model/models.py:7, never called).  edge_weight [E]: the message is
    fn.u_mul_e(h, w) instead of fn.copy_u (DGL GraphConv, `edge_weight` argument)."""
    n = feat.shape[0]
    src = torch.as_tensor(src, dtype=torch.int64)
    dst = torch.as_tensor(dst, dtype=torch.int64)
    in_feats, out_feats = (weight.shape if weight is not None else (0, 0))
    bs = lambda v: v.reshape((-1,) + (1,) * (feat.dim() - 1))
    if norm in ("left", "both"):
        d = torch.bincount(src, minlength=n).to(feat.dtype).clamp(min=1)
        feat = feat * bs(d.pow(-0.5) if norm == "both" else 1.0 / d)
    mult_first = weight is not None and in_feats > out_feats
    if mult_first:
        feat = feat @ weight
    msg = feat.index_select(0, src)
    if edge_weight is not None:
        msg = msg * torch.as_tensor(edge_weight, dtype=feat.dtype).reshape((-1,) + (1,) * (feat.dim() - 1))
    rst = torch.zeros_like(feat).index_add(0, dst, msg)
    if weight is not None and not mult_first:
        rst = rst @ weight
    if norm in ("right", "both"):
        d = torch.bincount(dst, minlength=n).to(feat.dtype).clamp(min=1)
        rst = rst * bs(d.pow(-0.5) if norm == "both" else 1.0 / d)
    if bias is not None:
        rst = rst + bias
    return rst


def gcn2conv(feat, feat_0, src, dst, weight1, bias, layer, alpha=0.1, lambda_=1.0, weight2=None, edge_weight=None):
    """dgl.nn.pytorch.GCN2Conv.forward (named by BASELINE.json north_star only).  edge_weight [E] as in GraphConv."""
    n = feat.shape[0]
    src = torch.as_tensor(src, dtype=torch.int64)
    dst = torch.as_tensor(dst, dtype=torch.int64)
    beta = math.log(lambda_ / layer + 1)
    norm = torch.bincount(dst, minlength=n).to(feat.dtype).clamp(min=1).pow(-0.5).unsqueeze(1)
    f = feat * norm
    msg = f.index_select(0, src)
    if edge_weight is not None:
        msg = msg * torch.as_tensor(edge_weight, dtype=feat.dtype).reshape((-1,) + (1,) * (feat.dim() - 1))
    f = torch.zeros_like(f).index_add(0, dst, msg) * norm
    f = f * (1 - alpha)
    f0 = feat_0 * alpha
    if weight2 is None:
        r = f + f0
        rst = (1 - beta) * r + beta * (r @ weight1)
    else:
        rst = (1 - beta) * f + beta * (f @ weight1) + (1 - beta) * f0 + beta * (f0 @ weight2)
    if bias is not None:
        rst = rst + bias
    return rst


# --------------------------------------------------------------------------------------
# next-row 3: edge-feature producer                          dgl/utils.py:54-77
# --------------------------------------------------------------------------------------

def quat_to_so3(q):
    """dgl/utils.py:54-66 - column-major flattening of R(q), q = (x, y, z, w)."""
    x, y, z, w = q[0], q[1], q[2], q[3]
    return np.array([1 - 2 * y ** 2 - 2 * z ** 2, 2 * x * y + 2 * z * w, 2 * x * z - 2 * y * w,
                     2 * x * y - 2 * z * w, 1 - 2 * x ** 2 - 2 * z ** 2, 2 * y * z + 2 * x * w,
                     2 * x * z + 2 * y * w, 2 * y * z - 2 * x * w, 1 - 2 * x ** 2 - 2 * y ** 2])


def cal_relative_pose(p1, p2):
    """dgl/utils.py:69-77 - (t2 - t1) || first 6 entries of R(q2 * conj(q1))."""
    p1t, p2t, p1r, p2r = p1[:3], p2[:3], p1[3:], p2[3:]
    x1, y1, z1, w1 = -p1r[0], -p1r[1], -p1r[2], p1r[3]
    x2, y2, z2, w2 = p2r[0], p2r[1], p2r[2], p2r[3]
    rp = quat_to_so3([(w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2), (w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2),
                      (w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2), (w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2)])[:6]
    return np.concatenate((p2t - p1t, rp))


def relative_pose_edges(poses, k):
    """Edge features for one K-robot sample in reference edge order (dgl/dataloader.py:116-122).
    poses: [K, 7] (xyz + quaternion xyzw)."""
    out = []
    for i in range(k):
        for j in range(k):
            if i != j:
                out.append(cal_relative_pose(poses[i], poses[j]))
    return np.stack(out, 0)


# --------------------------------------------------------------------------------------
# synthetic inputs - shared by tests, smoke and bench
# --------------------------------------------------------------------------------------

def init_layer_params(f, compress_gcn=True, seed=1, dtype=torch.float32):
    """PyTorch-default init of edge_encoder (nn.Linear(9,F), nn.Linear(F,2F); models.py:146-151)
    and conv (nn.Conv2d(2F,F,1); models.py:164) - uniform(+-1/sqrt(fan_in))."""
    g = torch.Generator().manual_seed(seed)

    def u(shape, fan_in):
        b = 1.0 / math.sqrt(fan_in)
        return ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * b).to(dtype)

    p = dict(w1=u((f, 9), 9), b1=u((f,), 9), w2=u((2 * f, f), f), b2=u((2 * f,), f))
    if compress_gcn:
        p.update(wc=u((f, 2 * f), 2 * f), bc=u((f,), 2 * f))
    return p


def synthetic_batch(g, k, f, p=64, seed=1, dtype=torch.float32):
    """G identical K-robot graphs; h ~ relu(N(0,1)) [N,F,P]; pose [E,9] = dt ~ N(0,3^2) || so3(q)[:6]."""
    gen = torch.Generator().manual_seed(seed)
    s, d = fully_connected_edges(k)
    b = batch_graphs([s] * g, [d] * g, [k] * g)
    n, e = g * k, g * k * (k - 1)
    h = torch.relu(torch.randn((n, f, p), generator=gen, dtype=torch.float32)).to(dtype)
    q = torch.randn((e, 4), generator=gen, dtype=torch.float64)
    q = (q / q.norm(dim=1, keepdim=True)).numpy()
    rot = quat_to_so3(q.T)[:6].T if e else np.zeros((0, 6))
    dt = (torch.randn((e, 3), generator=gen, dtype=torch.float64) * 3).numpy()
    pose = torch.from_numpy(np.concatenate([dt, rot], 1)).to(torch.float32)
    dy = torch.randn((n, f, p), generator=gen, dtype=torch.float32).to(dtype)
    return dict(h=h, pose=pose, dy=dy, **b)
