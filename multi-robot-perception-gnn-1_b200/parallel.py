"""Data-parallel sharding of robot-graph batches.

Graphs are independent units: rank r owns graphs [r*G/R, (r+1)*G/R) with its own node/edge tensors and a locally
built CSR - no data-path exchange.  The only collective is ONE all-reduce (NCCL over NVLink/NVSwitch) per layer
bucket of the GNN-layer gradients, launched from autograd hooks as soon as a bucket is complete so the
gcn2/conv2 reduction overlaps the gcn1/conv1 backward.  The reference has no multi-GPU path for this model
(torch.nn.DataParallel cannot scatter a DGLGraph; dgl/training.py:324-325) - this is new functionality.
"""
import torch
import torch.distributed as dist


def shard_range(num_graphs, rank, world):
    """Contiguous graph range of `rank`; sizes differ by at most one."""
    base, rem = divmod(num_graphs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class GradSync:
    """Bucketed, overlapped all-reduce (mean) of a module's parameter gradients.

    buckets: list of lists of parameter-name prefixes; default one bucket per direct child holding parameters,
    ordered last-to-first (the order backward produces them): [gcn2+conv2], [gcn1+conv1].

    Every bucket is ONE persistent flat fp32 buffer and every parameter's `.grad` is a view into it: autograd accumulates
    straight into the bucket, the collective runs in place on the bucket, and nothing is concatenated or copied back
    (at ResNet-50 width the two buckets are 67 MB each).  Call `zero_grad()` instead of setting `.grad = None` between
    steps; a `.grad` that was replaced anyway is folded back into its view when its hook fires.
    wire_dtype=torch.bfloat16 sends a bf16 copy of the bucket over NVLink (half the bytes) and converts the reduced values
    back into the fp32 bucket; the default keeps the collective in fp32 (1-vs-N equality to fp32 rounding).  Either way
    every rank ends with bit-identical gradients (check_ranks_agree()).
    """

    def __init__(self, module, group=None, buckets=None, average=True, wire_dtype=None):
        self.group, self.average = group, average
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        if wire_dtype not in (None, torch.float32, torch.bfloat16):
            raise ValueError("wire_dtype must be None / torch.float32 / torch.bfloat16")
        self.wire_dtype = None if wire_dtype in (None, torch.float32) else wire_dtype
        named = [(n, p) for n, p in module.named_parameters() if p.requires_grad]
        if buckets is None:
            layers = {}
            for n, p in named:
                digits = "".join(ch for ch in n.split(".")[0] if ch.isdigit()) or "0"
                layers.setdefault(digits, []).append((n, p))
            buckets = [layers[k] for k in sorted(layers, reverse=True)]
        else:
            buckets = [[(n, p) for n, p in named if any(n.startswith(pre) for pre in b)] for b in buckets]
        self.buckets = [b for b in buckets if b]
        self._flat, self._views, self._wire = [], [], []
        for b in self.buckets:
            dev = b[0][1].device
            total = sum(p.numel() for _, p in b)
            flat = torch.zeros(total, dtype=torch.float32, device=dev)
            views, off = [], 0
            for _, p in b:
                if p.dtype != torch.float32:
                    raise TypeError("GradSync expects fp32 master parameters")
                v = flat[off:off + p.numel()].view_as(p)
                if p.grad is not None:
                    v.copy_(p.grad)
                p.grad = v
                views.append(v)
                off += p.numel()
            self._flat.append(flat)
            self._views.append(views)
            self._wire.append(torch.empty(total, dtype=self.wire_dtype, device=dev) if self.wire_dtype is not None else None)
        self._work = [None] * len(self.buckets)
        self._pending = [0] * len(self.buckets)
        self._hooks = []
        self.overlap = True      # False: hooks stay quiet (e.g. gradients accumulated over several backward passes); finish() reduces
        self.enabled = True      # False: no collective at all (A/B timing of the exposed all-reduce cost)

        for bi, b in enumerate(self.buckets):
            for pi, (_, p) in enumerate(b):
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(bi, pi)))
        self.reset()

    def reset(self):
        self._pending = [len(b) for b in self.buckets]
        self._work = [None] * len(self.buckets)

    def zero_grad(self):
        """Zero every bucket (one memset per bucket) and make sure each parameter's .grad is its bucket view."""
        for bi, b in enumerate(self.buckets):
            self._flat[bi].zero_()
            for (_, p), v in zip(b, self._views[bi]):
                if p.grad is not v:
                    p.grad = v

    def _adopt(self, bi, pi):
        """A .grad that autograd (or the user) replaced: fold it into the bucket view and re-attach the view."""
        p, v = self.buckets[bi][pi][1], self._views[bi][pi]
        if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
            v.copy_(p.grad)
            p.grad = v

    def _make_hook(self, bi, pi):
        def hook(_param):
            self._adopt(bi, pi)
            if not self.overlap:
                return
            self._pending[bi] -= 1
            if self._pending[bi] == 0:
                self._launch(bi)
        return hook

    def _launch(self, bi):
        if self.world == 1 or not self.enabled:
            return
        buf = self._flat[bi]
        if self._wire[bi] is not None:
            self._wire[bi].copy_(buf)
            buf = self._wire[bi]
        self._work[bi] = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish(self):
        """Wait (on the current stream) for every bucket; the reduced (averaged) gradients are then in place.  Call after backward()."""
        for bi, b in enumerate(self.buckets):
            for pi in range(len(b)):
                self._adopt(bi, pi)
            if self.world > 1 and self.enabled:
                if self._work[bi] is None:      # a bucket whose hooks did not all fire (unused parameter, overlap off)
                    self._launch(bi)
                self._work[bi].wait()
                flat = self._flat[bi]
                if self._wire[bi] is not None:
                    flat.copy_(self._wire[bi])
                if self.average:
                    flat.div_(self.world)
        self.reset()

    def checksums(self):
        """Per-bucket (sum, sum of absolute values) in fp64 - cheap fingerprints for a cross-rank equality check."""
        return torch.stack([torch.stack((f.double().sum(), f.double().abs().sum())) for f in self._flat])

    def check_ranks_agree(self):
        """All-gather the bucket fingerprints and compare them bit for bit: after finish() every rank must hold the same reduced
        gradients.  Returns (ok, max relative deviation)."""
        cs = self.checksums()
        if self.world == 1:
            return True, 0.0
        out = [torch.empty_like(cs) for _ in range(self.world)]
        dist.all_gather(out, cs, group=self.group)
        ref = out[0]
        dev = max(float(((o - ref).abs() / ref.abs().clamp(min=1e-300)).max()) for o in out)
        return all(torch.equal(o, ref) for o in out), dev

    def grad_bytes(self):
        return sum(p.numel() * 4 for b in self.buckets for _, p in b)

    def wire_bytes(self):
        return self.grad_bytes() // (2 if self.wire_dtype is not None else 1)

    def remove(self):
        for h in self._hooks:
            h.remove()


def bind_to_gpu_numa_node(device_index):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (sysfs), so that pinned host buffers allocated afterwards
    are local to the GPU's PCIe root: with one rank per GPU on a two-socket host, host<->device copies otherwise cross the
    socket interconnect.  Returns the node number, or None when the topology is not exposed (containers, single-node hosts)."""
    import os
    import torch
    try:
        pr = torch.cuda.get_device_properties(device_index)
        path = f"/sys/bus/pci/devices/{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None
