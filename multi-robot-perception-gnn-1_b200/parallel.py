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
    """

    def __init__(self, module, group=None, buckets=None, average=True):
        self.group, self.average = group, average
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
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
        self._flat = [None] * len(self.buckets)
        self._work = [None] * len(self.buckets)
        self._pending = [0] * len(self.buckets)
        self._hooks = []
        self.overlap = True      # False: hooks stay quiet (e.g. gradients accumulated over several backward passes); finish() reduces

        for bi, b in enumerate(self.buckets):
            for _, p in b:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(bi)))
        self.reset()

    def reset(self):
        self._pending = [len(b) for b in self.buckets]
        self._work = [None] * len(self.buckets)

    def _make_hook(self, bi):
        def hook(_param):
            if not self.overlap:
                return
            self._pending[bi] -= 1
            if self._pending[bi] == 0:
                self._launch(bi)
        return hook

    def _launch(self, bi):
        if self.world == 1:
            return
        grads = [p.grad for _, p in self.buckets[bi]]
        flat = torch.cat([g.reshape(-1) for g in grads])
        self._flat[bi] = flat
        self._work[bi] = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish(self):
        """Wait for every bucket and write the reduced (averaged) gradients back.  Call after backward()."""
        for bi, b in enumerate(self.buckets):
            if self.world > 1:
                if self._work[bi] is None:      # a bucket whose hooks did not all fire (unused parameter)
                    if all(p.grad is not None for _, p in b):
                        self._launch(bi)
                    else:
                        continue
                self._work[bi].wait()
                flat = self._flat[bi]
                if self.average:
                    flat.div_(self.world)
                off = 0
                for _, p in b:
                    n = p.grad.numel()
                    p.grad.copy_(flat[off:off + n].view_as(p.grad))
                    off += n
        self.reset()

    def grad_bytes(self):
        return sum(p.numel() * 4 for b in self.buckets for _, p in b)

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
