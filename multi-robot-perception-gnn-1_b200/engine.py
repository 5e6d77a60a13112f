"""Host-buffer front end of the GNN block: the call a user with HOST tensors makes.

`HostPipeline.step(h, pose, dy, y_out, dh_out)` runs fwd+bwd of a GNNBlock over a batch that lives in pinned host memory
and returns the block output and the input gradient to pinned host memory.  Robot graphs are independent, so the batch is
cut into chunks of whole graphs and the three phases are overlapped on three CUDA streams:

    copy-in stream :  H2D of chunk i+1 (h, pose, dy)          |  PCIe is full duplex: in and out copies run concurrently
    compute stream :  fwd + bwd of chunk i (librobotgnn)      |  parameter gradients accumulate across chunks (sum)
    copy-out stream:  D2H of chunk i-1 (y, dh)                |

The reference has no counterpart (its loop moves a whole DGL batch with data.to('cuda:0'), dgl/training.py:183, and keeps
everything else on the device); this exists so that end-to-end throughput with host-resident data is bounded by the PCIe
link, not by serialising copies and kernels.
"""
import weakref

import torch

from .graph import robot_batch


class HostPipeline:
    def __init__(self, block, num_graphs, robots, feat, dtype=torch.bfloat16, device="cuda", chunks=8, slots=3, hw=(8, 8), out_feat=None,
                 grad_sync=None):
        if num_graphs % chunks:
            raise ValueError("num_graphs must be a multiple of chunks")
        self.block, self.G, self.K, self.F, self.dtype = block, num_graphs, robots, feat, dtype
        self.dev = torch.device(device)
        self.chunks, self.slots = chunks, slots
        self.gc = num_graphs // chunks
        self.nc, self.ec = self.gc * robots, self.gc * robots * (robots - 1)
        self.hw = hw
        # data-parallel runs: a parallel.GradSync whose buckets hold the parameter gradients.  The gradients of all chunks accumulate
        # in the buckets (hooks quiet), and ONE all-reduce per bucket is enqueued on the compute stream behind the last chunk, so
        # consecutive steps keep streaming (wait=False) - the next step's first chunks copy in while the reduction runs.
        self.grad_sync = grad_sync
        out_feat = feat if out_feat is None else out_feat        # 2 * feat for a block without the compress conv (cat output)
        self.graph = robot_batch(self.gc, robots, device=self.dev)
        self.graph.index()
        mk = lambda *s, dt=dtype: torch.empty(s, dtype=dt, device=self.dev)
        self.h_d = [mk(self.nc, feat, *hw) for _ in range(slots)]
        self.dy_d = [mk(self.nc, out_feat, *hw) for _ in range(slots)]
        self.pose_d = [mk(self.ec, 9, dt=torch.float32) for _ in range(slots)]
        self.s_in, self.s_comp, self.s_out = (torch.cuda.Stream(self.dev) for _ in range(3))
        self._slot_done = [None] * slots            # last compute event that read each input slot (carried across steps)
        es, px = self.h_d[0].element_size(), hw[0] * hw[1]
        self.h2d_bytes = num_graphs * robots * (feat + out_feat) * px * es + num_graphs * robots * (robots - 1) * 36
        self.d2h_bytes = num_graphs * robots * (feat + out_feat) * px * es

    def step(self, h, pose, dy, y_out, dh_out, wait=True):
        """h, dy, y_out, dh_out: pinned host [N, F, 8, 8]; pose: pinned host [E, 9] fp32.  Parameter .grad holds the sum over
        the whole batch on return (after the caller synchronises, e.g. torch.cuda.current_stream().synchronize()).
        wait=False leaves the calling stream untouched, so back-to-back steps stream through the three pipeline streams without
        a fill/drain bubble between them (the next step's H2D runs under this step's last D2H); call wait() before using the
        outputs or the gradients on the calling stream."""
        cur = torch.cuda.current_stream(self.dev)
        for s in (self.s_in, self.s_comp, self.s_out):
            s.wait_stream(cur)
        if self.grad_sync is not None:
            self.grad_sync.overlap = False
            with torch.cuda.stream(self.s_comp):
                self.grad_sync.zero_grad()
        else:
            for p in self.block.parameters():       # (allocated on s_comp: the memory goes back to that stream's pool)
                p.grad = None
        ev_comp = [None] * self.chunks
        for i in range(self.chunks):
            sl = i % self.slots
            n0, n1, e0, e1 = i * self.nc, (i + 1) * self.nc, i * self.ec, (i + 1) * self.ec
            with torch.cuda.stream(self.s_in):
                prev = ev_comp[i - self.slots] if i >= self.slots else self._slot_done[sl]
                if prev is not None:
                    self.s_in.wait_event(prev)                         # the slot's previous consumer (this or the last step) is done
                self.h_d[sl].copy_(h[n0:n1], non_blocking=True)
                self.dy_d[sl].copy_(dy[n0:n1], non_blocking=True)
                self.pose_d[sl].copy_(pose[e0:e1], non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(self.s_in)
            with torch.cuda.stream(self.s_comp):
                self.s_comp.wait_event(ev_in)
                hd = self.h_d[sl].detach().requires_grad_()
                y = self.block(self.graph, hd, self.pose_d[sl])
                y.backward(self.dy_d[sl])
                ev_comp[i] = torch.cuda.Event()
                ev_comp[i].record(self.s_comp)
                yd, dhd = y.detach(), hd.grad
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev_comp[i])
                yd.record_stream(self.s_out)
                dhd.record_stream(self.s_out)
                y_out[n0:n1].copy_(yd, non_blocking=True)
                dh_out[n0:n1].copy_(dhd, non_blocking=True)
        if self.grad_sync is not None:
            with torch.cuda.stream(self.s_comp):
                self.grad_sync.finish()             # stream-ordered: the host does not block, s_comp waits for NCCL
        for i in range(max(0, self.chunks - self.slots), self.chunks):
            self._slot_done[i % self.slots] = ev_comp[i]
        if wait:
            self.wait()

    def wait(self):
        """Make the calling stream wait for every copy and kernel the pipeline has enqueued."""
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_stream(self.s_out)
        cur.wait_stream(self.s_comp)


class GraphedStep:
    """fwd+bwd of a GNN block captured once into a CUDA graph and replayed per step.

    At the reference's script shape (8 graphs of 5 robots, dgl/training.py:28) one step is ~40 short kernels and the cost is
    host launch latency, not the GPU; a replayed graph removes it.  Every kernel of librobotgnn is capture-safe (no host
    sync, no allocation, tensor maps are by-value kernel parameters).  Inputs are copied into static buffers; outputs
    (`y`, `dh`) and the parameters' `.grad` are static tensors overwritten by each replay."""

    def __init__(self, block, graph, h, pose, dy, warmup=3):
        self.block, self.g = block, graph
        graph.index()
        self.h = h.detach().clone().requires_grad_()
        self.pose, self.dy = pose.detach().clone(), dy.detach().clone()
        cur = torch.cuda.current_stream(h.device)
        side = torch.cuda.Stream(h.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):          # first calls set kernel attributes / build caches: must happen outside capture
                self._run()
        cur.wait_stream(side)
        self.h.grad = None
        for p in block.parameters():
            p.grad = None
        self.cuda_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.cuda_graph):
            self._run()
        self.dh = self.h.grad
        self.param_grads = [(p, p.grad) for p in block.parameters()]     # static tensors every replay overwrites

    def _run(self):
        self.y = self.block(self.g, self.h, self.pose)
        self.y.backward(self.dy)

    def __call__(self, h=None, pose=None, dy=None):
        with torch.no_grad():
            if h is not None:
                self.h.copy_(h)
            if pose is not None:
                self.pose.copy_(pose)
            if dy is not None:
                self.dy.copy_(dy)
        self.cuda_graph.replay()
        for p, gr in self.param_grads:      # re-attach in case the caller replaced or cleared .grad (optimizer.zero_grad)
            p.grad = gr
        return self.y, self.dh


class GraphedForward:
    """Forward-only GNN slice (dgl/model/models.py:180-189) captured once into a CUDA graph and replayed per batch - the eval path
    (dgl/eval.py:23 runs batch 1 under torch.no_grad(): 5 robots, ~10 short kernels, pure launch latency when issued one by one).
    Valid for batches of `num_graphs` complete `robots`-node graphs in the reference's edge order (every batch of the reference's
    dataset is one; dataloader.py:88-95), so the topology is a constant of the captured graph and only h / pose are copied in.
    The weight casts are part of the captured graph: a replay always uses the module's current weights.
    The returned tensor is a static buffer that the next call overwrites."""

    def __init__(self, mod, num_graphs, robots, feat, dtype=torch.float32, device="cuda", hw=(8, 8), warmup=2):
        from .layers import gnn_slice
        self.dev = torch.device(device)
        self.g = robot_batch(num_graphs, robots, device=self.dev)
        self.g.index()
        n, e = num_graphs * robots, num_graphs * robots * (robots - 1)
        self.h = torch.zeros((n, feat) + tuple(hw), dtype=dtype, device=self.dev)
        self.pose = torch.zeros((e, 9), dtype=torch.float32, device=self.dev)
        cur = torch.cuda.current_stream(self.dev)
        side = torch.cuda.Stream(self.dev)
        side.wait_stream(cur)
        with torch.no_grad():
            with torch.cuda.stream(side):
                for _ in range(warmup):      # first calls set kernel attributes: must happen outside capture
                    gnn_slice(mod, self.g, self.h, self.pose)
            cur.wait_stream(side)
            self.cuda_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.cuda_graph):
                self.y = gnn_slice(mod, self.g, self.h, self.pose)

    def __call__(self, h, pose):
        with torch.no_grad():
            self.h.copy_(h)
            self.pose.copy_(pose)
        self.cuda_graph.replay()
        return self.y


class _SliceModule(torch.nn.Module):
    """gnn_slice over a fixed topology as a tensor-in / tensor-out module (what torch.cuda.make_graphed_callables captures)."""

    def __init__(self, mod, graph):
        super().__init__()
        self.mod, self.g = mod, graph

    def forward(self, h, pose):
        from .layers import gnn_slice
        return gnn_slice(self.mod, self.g, h, pose)


class GraphedTrainSlice:
    """The GNN slice as TWO captured CUDA graphs (forward, backward) behind one autograd node, so it sits inside an ordinary
    training loop (loss.backward(), optimizer.step(); dgl/training.py:192-216) and still costs two graph launches per step instead
    of ~26 kernel launches - at the reference's batch (8 graphs, training.py:28) the step is launch-bound.  Built with
    torch.cuda.make_graphed_callables over librobotgnn's capture-safe launch sequences; parameter gradients flow through
    autograd as usual (accumulated into .grad).  Fixed batch shape and topology (complete K-robot graphs)."""

    def __init__(self, mod, num_graphs, robots, feat, dtype=torch.float32, device="cuda", hw=(8, 8), h_requires_grad=True):
        self.dev = torch.device(device)
        self.g = robot_batch(num_graphs, robots, device=self.dev)
        self.g.index()
        n, e = num_graphs * robots, num_graphs * robots * (robots - 1)
        sample_h = torch.zeros((n, feat) + tuple(hw), dtype=dtype, device=self.dev, requires_grad=h_requires_grad)
        sample_pose = torch.zeros((e, 9), dtype=torch.float32, device=self.dev)
        self.fn = torch.cuda.make_graphed_callables(_SliceModule(mod, self.g), (sample_h, sample_pose))

    def __call__(self, h, pose):
        return self.fn(h, pose)


_GRAPHED = weakref.WeakKeyDictionary()      # module -> {(G, K, F, hw, dtype, device): GraphedForward}; never pickled with the module
GRAPHED_EVAL_MAX_NODES = 4096               # above this the kernels are long enough that launch latency does not matter


import os as _os
# gnn_slice_auto may also graph the TRAINING step (forward + backward graphs).  Opt-in (ROBOT_GNN_GRAPHED_TRAIN=1 or set this flag):
# capture happens inside the first training forward, and stream capture is process-global - a DataLoader pin-memory thread or an
# older autograd graph that is still alive can invalidate it (torch.cuda.make_graphed_callables' documented constraints).  The
# eval graph (no backward) has no such constraint and is always on.
GRAPHED_TRAIN = _os.environ.get("ROBOT_GNN_GRAPHED_TRAIN", "0") == "1"


def gnn_slice_auto(mod, g, h, pose=None, compute_dtype=None):
    """gnn_slice, replayed from captured CUDA graphs when that is safe and pays: a small CUDA batch of complete K-robot graphs
    (2..8 robots), intended return key.  Gradients disabled (eval): one forward graph (GraphedForward).  Gradients enabled
    (training): a forward and a backward graph behind one autograd node (GraphedTrainSlice).  Anything else runs the eager path."""
    from .layers import gnn_slice
    opt = mod.opt
    k = g.complete_k
    train = torch.is_grad_enabled()
    ok = (h.is_cuda and 2 <= k <= 8 and g.num_nodes() == g.batch_size * k and h.dim() == 4 and h.is_contiguous()
          and (compute_dtype is None or compute_dtype == h.dtype)
          and g.num_nodes() <= GRAPHED_EVAL_MAX_NODES and getattr(opt, "compress_gcn", False)
          and mod.gcn1.return_key == "images" and mod.gcn1.film and h.dtype in (torch.float32, torch.bfloat16)
          and not torch.cuda.is_current_stream_capturing() and (GRAPHED_TRAIN or not train))
    if not ok:
        return gnn_slice(mod, g, h, pose, compute_dtype=compute_dtype)
    pose = g.edata['pose'] if pose is None else pose
    key = (train, bool(h.requires_grad), g.batch_size, k, h.shape[1], tuple(h.shape[2:]), h.dtype, h.device)
    cache = _GRAPHED.setdefault(mod, {})
    gf = cache.get(key)
    if gf is None:
        if len(cache) >= 8:                 # a handful of batch shapes at most (last partial batch of an epoch)
            cache.clear()
        if train:
            gf = GraphedTrainSlice(mod, g.batch_size, k, h.shape[1], h.dtype, h.device, tuple(h.shape[2:]), h_requires_grad=h.requires_grad)
        else:
            gf = GraphedForward(mod, g.batch_size, k, h.shape[1], h.dtype, h.device, tuple(h.shape[2:]))
        cache[key] = gf
    return gf(h, pose)
