"""Batched robot graphs: the slice of DGL the reference's hot path touches, re-hosted on librobotgnn.

Replaces (reference file:line under /root/reference):
  dgl.graph((src, dst))       dgl/dataloader.py:99
  dgl.batch(graphs)           dgl/training.py:57-58, dgl/eval.py:54-55
  DGLGraph.ndata/.edata/.local_scope()/.to()/.in_degrees()/...   dgl/model/models.py:174-180,220-223
The index work itself (id shifting, degrees, in-/out-CSR, degree norms) runs in the CUDA kernels of
csrc/csr.cu; this module only keeps host-side bookkeeping (sizes, frames, caching).
"""
from contextlib import contextmanager
from types import SimpleNamespace

import torch

from . import _lib


class RobotGraph:
    """Duck-types the DGLGraph subset used by dgl/model/models.py and dgl/training.py.

    Edges are stored as per-graph LOCAL ids + (batch_num_nodes, batch_num_edges); global ids and the
    CSR index are produced on the device on first use and cached.
    """

    def __init__(self, local_src, local_dst, batch_num_nodes, batch_num_edges, complete_k=0):
        self._lsrc = local_src.to(torch.int64).reshape(-1)
        self._ldst = local_dst.to(torch.int64).reshape(-1)
        self._bnn = [int(v) for v in batch_num_nodes]
        self._bne = [int(v) for v in batch_num_edges]
        if sum(self._bne) != self._lsrc.numel() or self._lsrc.numel() != self._ldst.numel():
            raise ValueError("edge arrays do not match batch_num_edges")
        self._n, self._e = sum(self._bnn), sum(self._bne)
        self.ndata, self.edata = {}, {}
        # K if every graph is the reference's complete K-robot topology in its edge order (dataloader.py:88-95), else 0.
        # Established on the host where the edge lists are host data (graph(), robot_batch()); a claim on device data
        # is verified on the device by rgnn_check_complete and reported by validate().
        # complete_k may also be a per-graph list (batch() of graphs with different robot counts - a sample that lost a camera):
        # complete_ks keeps it, and a RAGGED batch whose graphs are all complete is bucketed by K for the compiled-topology kernels.
        if isinstance(complete_k, (list, tuple)):
            ks = [int(k) for k in complete_k]
            if len(ks) != len(self._bnn):
                raise ValueError("per-graph complete_k list does not match the number of graphs")
            self.complete_ks = [k if (k >= 2 and k == n and e == k * (k - 1)) else 0 for k, n, e in zip(ks, self._bnn, self._bne)]
            complete_k = self.complete_ks[0] if self.complete_ks and all(k == self.complete_ks[0] for k in self.complete_ks) else 0
        else:
            ck = int(complete_k) if complete_k else 0
            self.complete_ks = [ck if (ck >= 2 and n == ck and e == ck * (ck - 1)) else 0 for n, e in zip(self._bnn, self._bne)]
        self.complete_k = int(complete_k) if complete_k and all(n == complete_k for n in self._bnn) else 0
        self._global = None   # (src, dst, node_offsets, edge_offsets) on self.device
        self._index = None

    # ---- structure ---------------------------------------------------------------------------
    @property
    def device(self):
        return self._lsrc.device

    def num_nodes(self):
        return self._n

    number_of_nodes = num_nodes

    def num_edges(self):
        return self._e

    number_of_edges = num_edges

    @property
    def batch_size(self):
        return len(self._bnn)

    def batch_num_nodes(self):
        return torch.tensor(self._bnn, dtype=torch.int64, device=self.device)

    def batch_num_edges(self):
        return torch.tensor(self._bne, dtype=torch.int64, device=self.device)

    def _globals(self):
        if self._global is None:
            g = len(self._bnn)
            dev = self.device
            bnn = torch.tensor(self._bnn, dtype=torch.int64)
            bne = torch.tensor(self._bne, dtype=torch.int64)
            if dev.type == "cuda":
                bnn, bne = bnn.to(dev), bne.to(dev)
                noff = torch.empty(g + 1, dtype=torch.int64, device=dev)
                eoff = torch.empty(g + 1, dtype=torch.int64, device=dev)
                src = torch.empty(self._e, dtype=torch.int64, device=dev)
                dst = torch.empty(self._e, dtype=torch.int64, device=dev)
                with torch.cuda.device(dev):
                    _lib.check(_lib.lib().rgnn_batch_graphs(_lib.ptr(self._lsrc), _lib.ptr(self._ldst), _lib.ptr(bnn), _lib.ptr(bne),
                                                            g, self._e, _lib.ptr(noff), _lib.ptr(eoff), _lib.ptr(src), _lib.ptr(dst),
                                                            _lib.stream()), "rgnn_batch_graphs")
            else:
                # host bookkeeping for graphs still living on the CPU (DataLoader collate); no feature math here
                noff = torch.zeros(g + 1, dtype=torch.int64)
                eoff = torch.zeros(g + 1, dtype=torch.int64)
                noff[1:] = torch.cumsum(bnn, 0)
                eoff[1:] = torch.cumsum(bne, 0)
                shift = torch.repeat_interleave(noff[:-1], bne)
                src, dst = self._lsrc + shift, self._ldst + shift
            self._global = (src, dst, noff, eoff)
        return self._global

    def edges(self):
        s, d, _, _ = self._globals()
        return s, d

    def index(self):
        """Device CSR index (cached): int64 DGL-visible arrays + compact int32 copies + fp32 degree norms."""
        if self._index is None:
            _lib.require_cuda(self._lsrc)
            dev = self.device
            src, dst, noff, eoff = self._globals()
            n, e, g = self._n, self._e, len(self._bnn)
            i64 = lambda k: torch.empty(k, dtype=torch.int64, device=dev)
            ix = SimpleNamespace(N=n, E=e, G=g, complete_k=self.complete_k if 2 <= self.complete_k < 4096 else 0, max_nodes=max(self._bnn, default=0), max_edges=max(self._bne, default=0),
                                 src=src, dst=dst, node_offsets=noff, edge_offsets=eoff,
                                 in_deg=i64(n), out_deg=i64(n), in_ptr=i64(n + 1), in_src=i64(e), in_eid=i64(e),
                                 out_ptr=i64(n + 1), out_dst=i64(e), out_eid=i64(e),
                                 ptr32=torch.empty(2 * (n + 1) + g + 1, dtype=torch.int32, device=dev),
                                 idx32=torch.empty(max(6 * e, 1), dtype=torch.int32, device=dev),
                                 norm=torch.empty(max(3 * n, 1), dtype=torch.float32, device=dev),
                                 status=torch.zeros(1, dtype=torch.int32, device=dev))
            with torch.cuda.device(dev):
                _lib.check(_lib.lib().rgnn_build_batched_csr(
                    _lib.ptr(src), _lib.ptr(dst), n, e, _lib.ptr(noff), _lib.ptr(eoff), g,
                    _lib.ptr(ix.in_deg), _lib.ptr(ix.out_deg), _lib.ptr(ix.in_ptr), _lib.ptr(ix.in_src), _lib.ptr(ix.in_eid),
                    _lib.ptr(ix.out_ptr), _lib.ptr(ix.out_dst), _lib.ptr(ix.out_eid),
                    _lib.ptr(ix.ptr32), _lib.ptr(ix.idx32), _lib.ptr(ix.norm), _lib.ptr(ix.status), _lib.stream()),
                    "rgnn_build_batched_csr")
                if ix.complete_k:
                    _lib.check(_lib.lib().rgnn_check_complete(_lib.ptr(src), _lib.ptr(dst), g, ix.complete_k, e, _lib.ptr(ix.status),
                                                              _lib.stream()), "rgnn_check_complete")
            ix.in_ptr32, ix.out_ptr32 = ix.ptr32[: n + 1], ix.ptr32[n + 1: 2 * (n + 1)]
            ix.noff32 = ix.ptr32[2 * (n + 1):]
            ix.in_src32, ix.in_eid32 = ix.idx32[0:e], ix.idx32[e: 2 * e]
            ix.out_dst32, ix.out_eid32 = ix.idx32[2 * e: 3 * e], ix.idx32[3 * e: 4 * e]
            ix.src32, ix.dst32 = ix.idx32[4 * e: 5 * e], ix.idx32[5 * e: 6 * e]
            ix.inv_in_deg, ix.rsqrt_in, ix.rsqrt_out = ix.norm[0:n], ix.norm[n: 2 * n], ix.norm[2 * n: 3 * n]
            ix.buckets = self._k_buckets(dev) if not ix.complete_k else None
            self._index = ix
        return self._index

    def _k_buckets(self, dev):
        """Ragged batch of complete graphs -> [(K, graphs in the bucket, int32 [gb, 2] device tensor of (first node, first edge))],
        or None when some graph is not a (host-verified) complete 2..8-robot graph.  Host prefix sums + ONE small H2D copy at
        index-build time; the FiLM kernels then run once per bucket with the topology compiled in."""
        ks = self.complete_ks
        if not ks or any(k < 2 or k > 8 for k in ks) or self._n >= 2 ** 31 or self._e >= 2 ** 31:
            return None
        nb = eb = 0
        per = {}
        for k, n, e in zip(ks, self._bnn, self._bne):
            per.setdefault(k, []).extend((nb, eb))
            nb += n
            eb += e
        flat = torch.tensor([v for k in sorted(per) for v in per[k]], dtype=torch.int32)
        flat = flat.to(dev, non_blocking=False)
        out, off = [], 0
        for k in sorted(per):
            cnt = len(per[k]) // 2
            out.append((k, cnt, flat[off: off + 2 * cnt]))
            off += 2 * cnt
        return out

    def validate(self):
        """Host-synchronising check that no edge leaves its graph (DGL raises at construction time)."""
        st = int(self.index().status.item())
        if st & 1:
            raise ValueError("an edge references a node outside its graph")
        if st & 2:
            raise ValueError("graph was declared complete_k but its edge list is not the reference's complete-graph enumeration")
        return self

    def in_degrees(self):
        return self.index().in_deg

    def out_degrees(self):
        return self.index().out_deg

    def to(self, device, non_blocking=False):
        device = torch.device(device)
        if device.type == "cuda" and device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        if device == self.device:
            return self
        g = RobotGraph(self._lsrc.to(device, non_blocking=non_blocking), self._ldst.to(device, non_blocking=non_blocking),
                       self._bnn, self._bne, list(self.complete_ks))
        g.ndata = {k: v.to(device, non_blocking=non_blocking) for k, v in self.ndata.items()}
        g.edata = {k: v.to(device, non_blocking=non_blocking) for k, v in self.edata.items()}
        return g

    # ---- frames --------------------------------------------------------------------------------
    @contextmanager
    def local_scope(self):
        """Mutations of ndata/edata inside the scope are dropped on exit (dgl/model/models.py:174,220)."""
        old_n, old_e = self.ndata, self.edata
        self.ndata, self.edata = dict(old_n), dict(old_e)
        try:
            yield
        finally:
            self.ndata, self.edata = old_n, old_e

    def update_all(self, message_func, reduce_func):
        """Builtin message passing only: update_all(copy_u(u, m), sum|mean(m, out)) (dgl/model/dgl_models.py:126).
        Python UDFs (edge_udf / node_udf of models.py:207-211) are replaced wholesale by the fused GCN layer."""
        from . import functional as Fn
        if not (isinstance(message_func, _CopyU) and isinstance(reduce_func, _Reduce) and message_func.out == reduce_func.msg):
            raise NotImplementedError("RobotGraph.update_all supports the builtins copy_u + sum/mean; "
                                      "the FiLM UDF pair is fused inside robot_gnn GCN")
        x = self.ndata[message_func.u]
        self.ndata[reduce_func.out] = Fn.copy_u_reduce(self, x, reduce_func.op)


class _CopyU:
    def __init__(self, u, out):
        self.u, self.out = u, out


class _Reduce:
    def __init__(self, op, msg, out):
        self.op, self.msg, self.out = op, msg, out


class function:  # namespace mirroring dgl.function (dgl/model/models.py:8)
    @staticmethod
    def copy_u(u, out):
        return _CopyU(u, out)

    copy_src = copy_u

    @staticmethod
    def sum(msg, out):
        return _Reduce("sum", msg, out)

    @staticmethod
    def mean(msg, out):
        return _Reduce("mean", msg, out)


def graph(data, num_nodes=None, device=None, **_):
    """dgl.graph((src, dst)) - dgl/dataloader.py:99.  idtype int64, num_nodes = max id + 1, edge id = list position."""
    src, dst = data
    src = torch.as_tensor(src, dtype=torch.int64, device=device).reshape(-1)
    dst = torch.as_tensor(dst, dtype=torch.int64, device=device).reshape(-1)
    if src.shape != dst.shape:
        raise ValueError("src and dst must have the same length")
    if num_nodes is None:
        num_nodes = int(max(src.max().item(), dst.max().item())) + 1 if src.numel() else 0
    elif src.numel() and not src.is_cuda and int(max(src.max().item(), dst.max().item())) >= num_nodes:
        raise ValueError("edge endpoint >= num_nodes")
    return RobotGraph(src, dst, [num_nodes], [src.numel()], _host_complete_k(src, dst, num_nodes))


def _host_complete_k(src, dst, k):
    """K if the (host) edge list is the reference's complete-graph enumeration, else 0.  Device lists are not inspected."""
    if src.is_cuda or k < 2 or src.numel() != k * (k - 1):
        return 0
    ii = torch.arange(k).repeat_interleave(k)
    jj = torch.arange(k).repeat(k)
    keep = ii != jj
    return k if torch.equal(src, ii[keep]) and torch.equal(dst, jj[keep]) else 0


def complete_k_of(local_src, local_dst, batch_num_nodes):
    """K if EVERY graph of a batch given by host-side local edge lists is the reference's complete K-robot enumeration
    (dataloader.py:88-95), else 0 - for callers that rebuild a RobotGraph from stored arrays (graph caches)."""
    bnn = [int(v) for v in batch_num_nodes]
    if not bnn or len(set(bnn)) != 1 or bnn[0] < 2 or local_src.is_cuda:
        return 0
    k, g = bnn[0], len(bnn)
    e = k * (k - 1)
    src, dst = local_src.to(torch.int64).reshape(-1), local_dst.to(torch.int64).reshape(-1)
    if src.numel() != g * e or dst.numel() != g * e or not _host_complete_k(src[:e], dst[:e], k):
        return 0
    same = bool((src.view(g, e) == src[:e]).all()) and bool((dst.view(g, e) == dst[:e]).all())
    return k if same else 0


def batch(graphs):
    """dgl.batch - dgl/training.py:57-58: node ids of graph i shift by sum_{j<i} num_nodes_j, edges and every
    ndata/edata key concatenate in graph order."""
    graphs = list(graphs)
    if not graphs:
        raise ValueError("batch() needs at least one graph")
    keys_n, keys_e = set(graphs[0].ndata), set(graphs[0].edata)
    for g in graphs:
        if set(g.ndata) != keys_n or set(g.edata) != keys_e:
            raise ValueError("all graphs must carry the same ndata/edata keys")
    bg = RobotGraph(torch.cat([g._lsrc for g in graphs]), torch.cat([g._ldst for g in graphs]),
                    [n for g in graphs for n in g._bnn], [e for g in graphs for e in g._bne],
                    [k for g in graphs for k in g.complete_ks])
    for k in graphs[0].ndata:
        bg.ndata[k] = torch.cat([g.ndata[k] for g in graphs], 0)
    for k in graphs[0].edata:
        bg.edata[k] = torch.cat([g.edata[k] for g in graphs], 0)
    return bg


def robot_batch(num_graphs, num_robots, device=None):
    """G fully-connected K-robot graphs without self loops in the reference's src-major edge order
    (dgl/dataloader.py:88-95): eid(i->j) = i*(K-1) + (j if j < i else j-1)."""
    k = int(num_robots)
    ii = torch.arange(k).repeat_interleave(k)
    jj = torch.arange(k).repeat(k)
    keep = ii != jj
    s, d = ii[keep].to(torch.int64), jj[keep].to(torch.int64)
    g = RobotGraph(s.repeat(num_graphs).to(device) if device else s.repeat(num_graphs),
                   d.repeat(num_graphs).to(device) if device else d.repeat(num_graphs),
                   [k] * num_graphs, [k * (k - 1)] * num_graphs, complete_k=k)
    return g


def relative_pose_edges(poses):
    """Edge features for G samples of K robots from raw poses [G, K, 7] (xyz + quaternion xyzw) in the reference's edge order -
    the batched, on-device replacement of the per-sample cal_relative_pose loops (dgl/dataloader.py:116-122, dgl/utils.py:54-77).
    Returns [G*K*(K-1), 9] fp32, bit-identical to the reference's float64 computation rounded with .float()."""
    _lib.require_cuda(poses)
    if poses.dim() != 3 or poses.shape[-1] != 7 or poses.dtype not in (torch.float32, torch.float64):
        raise ValueError("poses must be [G, K, 7] float32/float64")
    poses = poses.contiguous()
    g, k = poses.shape[0], poses.shape[1]
    out = torch.empty((g * k * (k - 1), 9), dtype=torch.float32, device=poses.device)
    with torch.cuda.device(poses.device):
        _lib.check(_lib.lib().rgnn_relative_pose(int(poses.dtype == torch.float64), _lib.ptr(poses), _lib.ptr(out), g, k, _lib.stream()),
                   "rgnn_relative_pose")
    return out
