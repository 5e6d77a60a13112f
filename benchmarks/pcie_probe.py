"""Host <-> device copy rates the e2e number is bounded by: pinned H2D alone, D2H alone and both directions at once, with the
process unbound and bound to the GPU's NUMA node (pinned buffers allocated after binding).  One JSON line per setting."""
import json
import os
import subprocess
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def rates(nbytes=1 << 30, reps=5):
    dev = torch.device("cuda:0")
    hi = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    ho = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    hi.fill_(1)
    di, do = torch.empty(nbytes, dtype=torch.uint8, device=dev), torch.ones(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_event(e0); s2.wait_event(e0)
        for _ in range(reps):
            fn()
        cur = torch.cuda.current_stream()
        cur.wait_stream(s1); cur.wait_stream(s2)
        e1.record()
        torch.cuda.synchronize()
        return nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9

    def h2d():
        with torch.cuda.stream(s1):
            di.copy_(hi, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            ho.copy_(do, non_blocking=True)

    def both():
        h2d(); d2h()
    return {"h2d_GBps": round(timed(h2d), 1), "d2h_GBps": round(timed(d2h), 1), "duplex_each_GBps": round(timed(both), 1)}


if __name__ == "__main__":
    if len(sys.argv) > 1:
        from robot_gnn_b200 import parallel
        node = parallel.bind_to_gpu_numa_node(0) if sys.argv[1] == "bound" else None
        print(json.dumps({"setting": sys.argv[1], "numa_node": node, "cpus": len(os.sched_getaffinity(0)), **rates()}))
    else:
        for s in ("unbound", "bound"):
            subprocess.run([sys.executable, __file__, s])
