#!/usr/bin/env python
"""bench.py - robot-graph GNN-layer fwd+bwd throughput (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A step = one fwd+bwd pass of the GNN block (gcn1 -> compress -> gcn2 -> compress; dgl/model/models.py:180-189) over one
synthetic batch of K-robot graphs per GPU (+ the NCCL all-reduce of the GNN-layer gradients when N > 1).
Workload (BASELINE.json configs[1]): warehouse "gcn2" stack, ResNet-18 width F = 512, K = 5 robots,
P = 8x8, two layers + compress, bf16 I/O with fp32 accumulation, G = 4096 graphs per GPU (h = 1.34 GB >> 126 MB L2).
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

F_DIM, K_ROBOTS, PIX = 512, 5, 64
METRIC = "gnn_layer_fwd_bwd_robot_graphs_per_sec"
MEM_KERNELS = ("aggregate", "film_", "bias_grad", "mean_mix")      # HBM-bound kernel families (the rest are tensor-pipe GEMMs)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--graphs", type=int, default=4096, help="graphs per GPU per step")
    ap.add_argument("--feat", type=int, default=F_DIM)
    ap.add_argument("--robots", type=int, default=K_ROBOTS)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cuda-graph", action="store_true", help="replay the step from a captured CUDA graph (small batches)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=16, help="graph chunks the host-buffer pipeline overlaps")
    ap.add_argument("--cpu-graphs", type=int, default=192, help="graphs in the bounded CPU-baseline sample")
    ap.add_argument("--stack", default="film", choices=["film", "gcnii"],
                    help="film = the reference's GNN block (FiLM-GCN layers + compress, dgl/model/models.py:180-189); gcnii = the north star's "
                         "GCN2Conv stack (DGL formula, SURVEY App. B) on channels-last [N, 64, F] features")
    ap.add_argument("--layers", type=int, default=2, choices=[1, 2], help="1 = gcn1 only (-multi_gcn off)")
    ap.add_argument("--no-compress", action="store_true", help="no 1x1 compress conv (-compress_gcn off; single layer only)")
    ap.add_argument("--no-other-configs", action="store_true",
                    help="skip the other_configs block (configs[3] at ResNet-50 width strong- and weak-scaled, one K=64 cell of configs[4])")
    ap.add_argument("--wire", default="fp32", choices=["fp32", "bf16"], help="gradient all-reduce wire format (fp32 master buckets either way)")
    ap.add_argument("--torch-baseline", action="store_true",
                    help="also time a plain PyTorch-on-GPU restatement (index_select / index_add_ / conv2d) of the same step")
    return ap.parse_args()


def workload_name(a):
    if a.stack == "gcnii":
        return (f"GCNII stack ({a.layers}x GCN2Conv: D^-1/2 A D^-1/2 propagation, initial residual alpha=0.1, identity mapping), F={a.feat}, "
                f"K={a.robots} robots, channels-last [N, 64, F], {a.graphs} graphs/GPU, {a.dtype}")
    width = {512: " (resnet18 width)", 1280: " (mobilenetv2 width)", 2048: " (resnet50 width)"}.get(a.feat, "")
    if a.layers == 2:
        stack = "warehouse gcn2 stack (2x FiLM-GCN + 1x1 compress)"
    else:
        stack = "gcn stack (1x FiLM-GCN" + (")" if a.no_compress else " + 1x1 compress)")
    return f"{stack}, F={a.feat}{width}, K={a.robots} robots, P=8x8, {a.graphs} graphs/GPU, {a.dtype}"


def torch_gpu_step_fn(blk, g, a):
    """A plain PyTorch-on-GPU restatement of the same block (what a user without DGL or this library would write): gather,
    broadcast FiLM, index_add_ mean, cat, conv2d.  Uses the block's own parameters; timed as a reported baseline only."""
    import torch.nn.functional as Fn
    src, dst = g.edges()
    src, dst = src.long(), dst.long()
    inv = 1.0 / g.in_degrees().clamp(min=1).to(torch.float32)

    def layer(gcn, conv, h, pose):
        l0, l2 = gcn.edge_encoder.layers[0], gcn.edge_encoder.layers[2]
        z = torch.sigmoid(Fn.linear(torch.relu(Fn.linear(pose.to(h.dtype), l0.weight.to(h.dtype), l0.bias.to(h.dtype))),
                                    l2.weight.to(h.dtype), l2.bias.to(h.dtype)))
        m = z[:, 0::2, None, None] * h.index_select(0, src) + z[:, 1::2, None, None]
        agg = torch.zeros_like(h).index_add_(0, dst, m) * inv.to(h.dtype)[:, None, None, None]
        if conv is None:
            return torch.cat((h, agg), 1)
        return Fn.conv2d(torch.cat((h, agg), 1), conv.weight.to(h.dtype), conv.bias.to(h.dtype))

    if a.stack == "gcnii":
        k = a.robots
        a_hat = ((torch.ones(k, k, device=src.device) - torch.eye(k, device=src.device)) / (k - 1))

        def fwd_gcnii(h, pose):
            x0 = h.reshape(h.shape[0], 64, -1)
            x = x0
            for l in blk.layers:
                agg = torch.einsum("ab,gbpf->gapf", a_hat.to(x.dtype), x.reshape(-1, k, 64, x.shape[-1])).reshape(x.shape)
                r = (1 - l.alpha) * agg + l.alpha * x0
                x = (1 - l.beta) * r + l.beta * (r @ l.weight1.to(x.dtype)) + l.bias.to(x.dtype)
            return x.reshape(h.shape)
        return fwd_gcnii

    def fwd(h, pose):
        y = layer(blk.gcn1, None if a.no_compress else blk.conv1, h, pose)
        if a.layers == 2:
            y = layer(blk.gcn2, blk.conv2, y, pose)
        return y
    return fwd


# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def config_of(a, world):
    """The `config` object of the JSON line - identical for both arms (`--impl ours|reference`) of one command line."""
    return {"workload": workload_name(a), "graphs_per_gpu": a.graphs, "robots": a.robots, "feature_dim": a.feat, "pixels": PIX,
            "stack": a.stack, "layers": a.layers, "compress": (not a.no_compress) if a.stack == "film" else None,
            "l2_policy": "inputs (1.3 GB/GPU at the default workload) larger than the 126 MB L2",
            "parallelism": f"dp{world}: graphs sharded, NCCL all-reduce of GNN-layer grads only"}


def cpu_reference_rate(a, reps):
    """The reference's CPU path on a bounded sample of the same workload, fp32, all host threads.  Preferred: the reference's own
    classes, byte-compiled into oracle/_ref by build() (kind "reference"); else the op-for-op port oracle/ref_port.py (kind "port").
    Both run over oracle/dgl_shim (DGL is not installable).  -> (graphs/s, cores, seconds per pass, kind)"""
    from oracle import ref_port, ref_runner
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if a.stack == "gcnii":
        return a.cpu_graphs / ref_port.time_gcnii(a.feat, a.robots, a.cpu_graphs, layers=a.layers, reps=reps, threads=cores), cores, None, "port"
    kind = "reference" if ref_runner.available() else "port"
    mod = ref_runner if kind == "reference" else ref_port
    sec = mod.time_block(a.feat, a.robots, a.cpu_graphs, not a.no_compress, a.layers == 2, reps=reps, threads=cores)
    return a.cpu_graphs / sec, cores, sec, kind


CPU_WHAT = {"reference": "unmodified reference classes (dgl/model/models.py byte-compiled into oracle/_ref: multi_view_dgl_model members, GCN, "
                         "edge_encoder, edge_udf, node_udf) over the DGL-semantics shim, GNN slice models.py:181-189",
            "port": "op-for-op port of the reference block (oracle/ref_port.py) over the DGL-semantics shim"}


def run_reference(a, rank):
    """`--impl reference`: the reference's own CPU implementation of the path on this box's host cores.  One "step" = one fwd+bwd
    pass over the bounded sample; `steps` / `warmup` report the passes actually run (bounded so the arm ends within minutes)."""
    if rank != 0:
        return
    t0 = time.time()
    reps = max(1, min(a.steps, 5))
    rate, cores, sec, kind = cpu_reference_rate(a, reps=reps)
    sec = a.cpu_graphs / rate if sec is None else sec
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "graphs/s", "n_gpus": a.gpus, "steps": reps,
            "warmup": 1, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config_of(a, a.gpus),
            "cpu_baseline": {"value": rate, "unit": "graphs/s", "cores": cores, "kind": kind,
                             "sample": f"{CPU_WHAT[kind]}; fwd+bwd of {a.cpu_graphs} graphs per step (bounded sample of the workload), "
                                       f"best of {reps} after 1 warm-up, torch {torch.__version__} CPU, {cores} threads"},
            "e2e": {"value": rate, "unit": "graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "requested": {"steps": a.steps, "warmup": a.warmup}, "wall_s": time.time() - t0}
    print(json.dumps(line))


TRAFFIC_FILE = "traffic_r02.json"


def peaks_for_other():
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pk = {}
    return pk.get("hbm_gbs", 6650.0), pk.get("bf16_tflops_sustained", 1400.0)


def allreduce_cost(sync, step, barrier, dist, dev, world, n, ms_step_with):
    """What the gradient all-reduce costs a step: (a) the collective alone, back to back on idle GPUs; (b) EXPOSED time = the step
    with the reduction minus the same step without it (sync.enabled = False), both timed over n steps, max over ranks."""
    def timed(fn):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def only_reduce():
        for bi in range(len(sync.buckets)):
            sync._launch(bi)
        sync.finish()
    for _ in range(2):
        only_reduce()
    ar_ms = timed(only_reduce)
    sync.enabled = False
    step()
    ms_without = timed(step)
    sync.enabled = True
    ms_with = timed(step)
    return {"grad_bytes": sync.grad_bytes(), "wire_bytes": sync.wire_bytes(), "buckets": len(sync.buckets),
            "allreduce_ms_alone": round(ar_ms, 4), "algbw_gbs": round(sync.wire_bytes() / (ar_ms * 1e-3) / 1e9, 1),
            "ms_per_step_with": round(ms_with, 4), "ms_per_step_without": round(ms_without, 4),
            "exposed_ms": round(ms_with - ms_without, 4), "exposed_frac": round((ms_with - ms_without) / ms_with, 4)}


def other_configs(R, dist, dev, rank, world, wire, peaks):
    """BASELINE.json configs[3] (ResNet-50 width F=2048, train_camera_warehouse_gcn2_resnet50.sh:4) strong- and weak-scaled, and one
    cell of configs[4] (K = 64 robots), measured in the same process group right after the headline workload: graphs/s, the
    dominant kernel's roofline, and (N > 1) the all-reduce cost with its exposed share.  Every entry is a full fwd+bwd(+all-reduce)."""
    import types
    hbm_peak, tc_peak = peaks
    out = []
    cells = [dict(name="C4 strong: F=2048, K=5, 2 layers + compress, bf16, 8192 graphs in total", F=2048, K=5, G=max(1, 8192 // world), scaling="strong", steps=4),
             dict(name="C4 weak: F=2048, K=5, 2 layers + compress, bf16, 1024 graphs per GPU", F=2048, K=5, G=1024, scaling="weak", steps=8),
             dict(name="C5 cell: F=512, K=64 robots, 2 layers + compress, bf16, 128 graphs per GPU", F=512, K=64, G=128, scaling="weak", steps=6),
             dict(name="ragged batch: robots per graph mixed over {3, 4, 5} (dropped cameras, dataloader.py:88-99), F=512, 2 layers + compress, "
                       "bf16, 4096 graphs per GPU", F=512, K=5, G=4096, scaling="weak", steps=8, ragged=(3, 4, 5))]
    for c in cells:
        F, K, G = c["F"], c["K"], c["G"]
        N, E = G * K, G * K * (K - 1)
        ks = None
        if c.get("ragged"):
            ks = [c["ragged"][i % len(c["ragged"])] for i in range(G)]
            N, E = sum(ks), sum(k * (k - 1) for k in ks)
        torch.manual_seed(1)
        blk = R.GNNBlock(types.SimpleNamespace(feature_dim=F, compress_gcn=True, multi_gcn=True)).to(dev)
        sync = R.parallel.GradSync(blk, wire_dtype=wire) if world > 1 else None
        gen = torch.Generator(device=dev).manual_seed(1 + rank)
        h = torch.relu(torch.randn((N, F, 8, 8), generator=gen, device=dev, dtype=torch.float32)).bfloat16().requires_grad_()
        dy = torch.randn((N, F, 8, 8), generator=gen, device=dev, dtype=torch.float32).bfloat16()
        pose = torch.randn((E, 9), generator=gen, device=dev, dtype=torch.float32)
        if ks is None:
            g = R.robot_batch(G, K, device=dev)
        else:       # per-sample dgl.graph + dgl.batch as the DataLoader collate does (training.py:57-58), then .to(device)
            tmpl = {k: R.robot_batch(1, k) for k in set(ks)}
            g = R.batch([tmpl[k] for k in ks]).to(dev)
        g.index()

        def step():
            h.grad = None
            if sync is not None:
                sync.zero_grad()
            else:
                for p in blk.parameters():
                    p.grad = None
            blk(g, h, pose).backward(dy)
            if sync is not None:
                sync.finish()

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
        for _ in range(3):
            step()
        barrier()
        R._lib.PROFILE = {}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(c["steps"]):
            step()
        e1.record()
        barrier()
        prof = R._lib.profile_summary()
        R._lib.PROFILE = None
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms_step = float(ms.item()) / c["steps"]
        top, d = max(prof.items(), key=lambda kv: kv[1]["ms_total"])
        mem = top.startswith(MEM_KERNELS)
        ach = d["work_total"] / (d["ms_total"] * 1e-3) / (1e9 if mem else 1e12)
        ent = {"name": c["name"], "scaling": c["scaling"], "n_gpus": world, "graphs_per_gpu": G, "robots": K, "feature_dim": F,
               "value": world * G / (ms_step * 1e-3), "unit": "graphs/s", "ms_per_step": round(ms_step, 4), "steps": c["steps"], "warmup": 3,
               "roofline": {"kernel": top, "bound": "hbm" if mem else "tensor", "achieved": round(ach, 2), "unit": "GB/s" if mem else "TFLOP/s",
                            "peak": hbm_peak if mem else tc_peak, "frac": round(ach / (hbm_peak if mem else tc_peak), 4),
                            "ms_avg_per_launch": round(d["ms_avg"], 4), "share": round(d["ms_total"] / (ms_step * c["steps"]), 4)}}
        ent["hbm_kernels"] = {n: {"ms_avg": round(v["ms_avg"], 4), "GB/s": round(v["work_total"] / (v["ms_total"] * 1e-3) / 1e9, 1),
                                  "frac": round(v["work_total"] / (v["ms_total"] * 1e-3) / 1e9 / hbm_peak, 4)}
                              for n, v in prof.items() if n.startswith(MEM_KERNELS) and v["ms_total"] > 0}
        if sync is not None:
            ok, dev_rel = sync.check_ranks_agree()
            ent["grad_sync_check"] = "ok" if ok else f"MISMATCH ({dev_rel:.3e})"
            ent["comm"] = allreduce_cost(sync, step, barrier, dist, dev, world, max(3, c["steps"] // 2), ms_step)
            sync.remove()
        out.append(ent)
        del blk, sync, h, dy, pose, g
        import gc
        gc.collect()
        torch.cuda.empty_cache()
    out.append(script_shape_latency(R, dev))
    return out


def script_shape_latency(R, dev):
    """The reference's REAL batch (dgl/training.py:28: 8 graphs of 5 robots; dgl/eval.py:23: 1 graph): launch-bound, so what is
    measured is host + launch latency of one step on one GPU - eager (kernel by kernel through the C ABI) and replayed from CUDA graphs
    (GraphedTrainSlice: forward + backward graphs behind one autograd node; GraphedForward: the eval slice).  Wall-clock per step over
    200 steps bracketed by synchronize (CUDA events cannot see host latency)."""
    import sys
    import types
    eng = sys.modules[R.GraphedForward.__module__]
    F, K = 512, 5
    torch.manual_seed(1)
    blk = R.GNNBlock(types.SimpleNamespace(feature_dim=F, compress_gcn=True, multi_gcn=True)).to(dev)
    res = {"name": "script shape: F=512, K=5, 2 layers + compress, bf16, G=8 (training.py:28) and G=1 (eval.py:23), one GPU", "unit": "ms/step"}
    n_steps = 200
    for G, mode in ((8, "train"), (8, "eval"), (1, "eval")):
        N, E = G * K, G * K * (K - 1)
        g = R.robot_batch(G, K, device=dev)
        g.index()
        h = torch.relu(torch.randn((N, F, 8, 8), device=dev)).bfloat16()
        dy = torch.randn((N, F, 8, 8), device=dev).bfloat16()
        g.edata["pose"] = torch.randn((E, 9), device=dev)
        for graphed in (False, True):
            old_flag, eng.GRAPHED_TRAIN = eng.GRAPHED_TRAIN, True
            fn = R.gnn_slice_auto if graphed else R.gnn_slice

            params = list(blk.parameters())          # (as an optimizer holds them; walking the module tree costs 70 us per step)

            def step():
                if mode == "eval":
                    with torch.no_grad():
                        return fn(blk, g, h)
                hh = h.detach().requires_grad_()
                for p in params:
                    p.grad = None
                fn(blk, g, hh).backward(dy)
            try:
                for _ in range(10):
                    step()
                torch.cuda.synchronize()
                n0 = R._lib.lib().rgnn_launch_count()
                t0 = time.perf_counter()
                for _ in range(n_steps):
                    step()
                torch.cuda.synchronize()
                ms = (time.perf_counter() - t0) / n_steps * 1e3
                res[f"G{G}_{mode}_{'graphed' if graphed else 'eager'}"] = {"ms_per_step": round(ms, 4), "graphs_per_s": round(G / (ms * 1e-3)),
                                                                       "c_abi_launches_per_step": (R._lib.lib().rgnn_launch_count() - n0) / n_steps}
            finally:
                eng.GRAPHED_TRAIN = old_flag
    return res


# ------------------------------------------------------------------------------------------------------------
def main():
    a = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if a.impl == "reference":
        run_reference(a, rank)
        return
    import torch.distributed as dist
    import types
    import robot_gnn_b200 as R           # the oracle is imported by the cpu_baseline leg only (cpu_reference_rate)

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = R.parallel.bind_to_gpu_numa_node(local) if world > 1 else None     # before any pinned host buffer exists
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    R._lib.check(R._lib.lib().rgnn_device_ok(), "rgnn_device_ok")

    dtype = torch.bfloat16 if a.dtype == "bf16" else torch.float32
    G, K, F = a.graphs, a.robots, a.feat
    N, E = G * K, G * K * (K - 1)
    torch.manual_seed(1)                 # reference default seed (training.py:24); PyTorch-default init of Linear / Conv2d
    if a.no_compress and a.layers == 2:
        raise SystemExit("two layers need the compress conv (dgl/model/models.py:169)")
    if a.stack == "gcnii":
        class GcniiStack(torch.nn.Module):
            """x_{l+1} = GCN2Conv_l(g, x_l, x_0) over channels-last [N, 64, F] views of the [N, F, 8, 8]-sized buffers."""
            def __init__(self):
                super().__init__()
                self.layers = torch.nn.ModuleList(R.GCN2Conv(F, layer=i + 1, alpha=0.1, lambda_=1) for i in range(a.layers))
                for l in self.layers:
                    torch.nn.init.normal_(l.weight1, std=F ** -0.5)

            def forward(self, g, h, pose=None):
                x0 = h.reshape(h.shape[0], 64, -1)
                x = x0
                for l in self.layers:
                    x = l(g, x, x0)
                return x.reshape(h.shape)
        blk = GcniiStack()
        a.no_compress = False
    else:
        blk = R.GNNBlock(types.SimpleNamespace(feature_dim=F, compress_gcn=not a.no_compress, multi_gcn=a.layers == 2))
    blk.to(dev)
    wire = torch.bfloat16 if a.wire == "bf16" else None
    sync = R.parallel.GradSync(blk, wire_dtype=wire) if world > 1 else None

    # synthetic inputs of the scripts' shape, seeded per rank; resident in HBM for `value`
    gen = torch.Generator(device=dev).manual_seed(1 + rank)
    h = torch.relu(torch.randn((N, F, 8, 8), generator=gen, device=dev, dtype=torch.float32)).to(dtype).requires_grad_()
    f_out = 2 * F if a.no_compress else F       # without the compress conv the block returns cat(h, agg) (models.py:182)
    dy = torch.randn((N, f_out, 8, 8), generator=gen, device=dev, dtype=torch.float32).to(dtype)
    pose = torch.randn((E, 9), generator=gen, device=dev, dtype=torch.float32)
    pose[:, :3] *= 3.0
    g = R.robot_batch(G, K, device=dev)
    g.index()

    def step():
        h.grad = None
        if sync is not None:
            sync.zero_grad()          # gradients are views into the persistent all-reduce buckets
        else:
            for p in blk.parameters():
                p.grad = None
        y = blk(g, h, pose)
        y.backward(dy)
        if sync is not None:
            sync.finish()
        return y

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if a.cuda_graph and world == 1:
        graphed = R.GraphedStep(blk, g, h, pose, dy)
        step = lambda: graphed()      # noqa: E731  (inputs already live in the static buffers)

    for _ in range(max(a.warmup, 3)):
        step()
    barrier()
    # ---- timed region: device-resident inputs ---------------------------------------------------------------
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    R._lib.PROFILE = {}
    launches0 = R._lib.lib().rgnn_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(a.steps):
        step()
    ev1.record()
    barrier()
    launches = R._lib.lib().rgnn_launch_count() - launches0
    prof = R._lib.profile_summary()
    R._lib.PROFILE = None
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clk = clocks.stop() if rank == 0 else None
    value = world * G * a.steps / (ms_total * 1e-3)

    # ---- N > 1: do the ranks agree on the reduced gradients, and what does the all-reduce cost the step? -------------------
    comm = None
    if sync is not None:
        ok, dev_rel = sync.check_ranks_agree()
        comm = {"grad_sync_check": "ok" if ok else f"MISMATCH (max relative deviation of bucket checksums {dev_rel:.3e})"}
        comm.update(allreduce_cost(sync, step, barrier, dist, dev, world, max(3, min(a.steps, 10)), ms_total / a.steps))

    # ---- e2e: the public call with HOST buffers; H2D of the step's inputs and D2H of its results inside the timed region
    e2e = None
    if not a.no_e2e:
        hh = h.detach().cpu().pin_memory(); hdy = dy.cpu().pin_memory(); hpose = pose.cpu().pin_memory()
        out_y = torch.empty((N, f_out, 8, 8), dtype=dtype).pin_memory(); out_dh = torch.empty((N, F, 8, 8), dtype=dtype).pin_memory()
        h2d = hh.numel() * hh.element_size() + hdy.numel() * hdy.element_size() + hpose.numel() * 4
        d2h = 2 * out_y.numel() * out_y.element_size()

        pipe = R.HostPipeline(blk, G, K, F, dtype=dtype, device=dev, chunks=a.e2e_chunks, out_feat=f_out, grad_sync=sync)
        h2d, d2h = pipe.h2d_bytes, pipe.d2h_bytes

        def e2e_step():
            # steps stream back to back through the pipeline (every step still copies its own inputs in and its own results out
            # inside the timed region).  N > 1: the gradients of all chunks accumulate in the all-reduce buckets and ONE reduction per
            # bucket is enqueued on the compute stream behind the last chunk - no per-step join, the next step's copies run under it
            pipe.step(hh, hpose, hdy, out_y, out_dh, wait=False)

        n_e2e = max(3, min(a.steps, 10))
        e2e_step(); e2e_step()
        pipe.wait()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_e2e):
            e2e_step()
        pipe.wait()
        e1.record()
        barrier()
        ems = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        e2e = {"value": world * G * n_e2e / (float(ems.item()) * 1e-3), "unit": "graphs/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": n_e2e,
               "api": f"HostPipeline.step: pinned host h/pose/dy in, y/dh out, {a.e2e_chunks} graph chunks overlapped on 3 streams"}

    torch_gpu = None
    if a.torch_baseline and rank == 0:
        fwd = torch_gpu_step_fn(blk, g, a)
        torch.backends.cudnn.allow_tf32 = False          # our fp32 mode is true fp32 (1e-5 parity); keep the baseline's arithmetic the same
        torch.backends.cuda.matmul.allow_tf32 = False

        def tstep():
            h.grad = None
            for p in blk.parameters():
                p.grad = None
            fwd(h, pose).backward(dy)
        for _ in range(3):
            tstep()
        torch.cuda.synchronize()
        nt = max(3, min(a.steps, 10))
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(nt):
            tstep()
        t1.record()
        torch.cuda.synchronize()
        torch_gpu = {"value": G * nt / (t0.elapsed_time(t1) * 1e-3), "unit": "graphs/s", "ms_per_step": t0.elapsed_time(t1) / nt,
                     "what": f"plain PyTorch {torch.__version__} eager on the same GPU and inputs (index_select / index_add_ / cuBLAS / cuDNN), 1 GPU"}

    # ---- the other BASELINE configs in the same run (every rank takes part; rank 0 reports) ---------------------------------
    other = None
    if not a.no_other_configs and a.stack == "film" and (a.feat, a.robots, a.graphs, a.dtype, a.layers) == (F_DIM, K_ROBOTS, 4096, "bf16", 2):
        del h, dy, pose, g, blk, sync
        if not a.no_e2e:
            del pipe, hh, hdy, hpose, out_y, out_dh
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        other = other_configs(R, dist, dev, rank, world, wire, peaks_for_other())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (live CUDA-event timings from the timed region) --------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    tc_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "measured" if peaks else "fallback"
    by_time = sorted(prof.items(), key=lambda kv: -kv[1]["ms_total"])
    kernels = {}
    for name, d in prof.items():
        is_mem = name.startswith(MEM_KERNELS)
        ach = d["work_total"] / (d["ms_total"] * 1e-3) / (1e9 if is_mem else 1e12) if d["ms_total"] > 0 else 0.0
        kernels[name] = {"calls": d["calls"], "ms_avg": round(d["ms_avg"], 4), "share": round(d["ms_total"] / ms_total, 4),
                         "achieved": round(ach, 2), "unit": "GB/s" if is_mem else "TFLOP/s",
                         "frac": round(ach / (hbm_peak if is_mem else tc_peak), 4)}
        if d.get("bytes_total", 0) > 0 and d["ms_total"] > 0:      # a fused kernel: its own algorithmic HBM bytes beside its GEMM's flops
            gbs = d["bytes_total"] / (d["ms_total"] * 1e-3) / 1e9
            kernels[name].update({"algorithmic_bytes_per_launch": d["bytes_total"] / max(1, d["calls"]), "hbm_GBps": round(gbs, 1),
                                  "hbm_frac": round(gbs / hbm_peak, 4)})
    if not by_time:      # CUDA-graph replay: no per-call events exist
        by_time = [("cuda_graph_replay", dict(work_total=0.0, calls=1))]
        kernels["cuda_graph_replay"] = {"calls": a.steps, "ms_avg": round(ms_total / a.steps, 4), "share": 1.0, "achieved": 0.0,
                                        "unit": "GB/s", "frac": 0.0}
    top, topd = by_time[0]
    top_mem = top.startswith(MEM_KERNELS)
    traffic, traffic_source = None, None
    try:    # measured once under `ncu --set full` for this exact workload and committed under profiles/ (NOT measured in this run)
        if (G, K, F, a.dtype, a.stack, a.layers) == (4096, 5, 512, "bf16", "film", 2) and not a.no_compress:
            tj = json.load(open(os.path.join(ROOT, "profiles", TRAFFIC_FILE)))
            traffic = tj.get(top)
            traffic_source = f"profiles/{TRAFFIC_FILE}: dram__bytes_read.sum + dram__bytes_write.sum per launch from {tj.get('_source', 'an ncu --set full capture')}"
    except Exception:
        pass
    roofline = {"kernel": top, "bound": "hbm" if top_mem else "tensor", "achieved": kernels[top]["achieved"],
                "peak": hbm_peak if top_mem else tc_peak, "unit": kernels[top]["unit"], "frac": kernels[top]["frac"],
                "traffic": traffic, "traffic_source": traffic_source, "algorithmic_per_launch": topd["work_total"] / max(1, topd["calls"]),
                "peak_source": peak_src, "ms_avg_per_launch": kernels[top]["ms_avg"]}
    if "hbm_frac" in kernels[top]:
        roofline["note"] = ("fused dX GEMM + FiLM backward: its MMA phase reaches 0.57-0.60 of the tensor peak (B operand staged as 64-byte rows so "
                            "that three pixel-half accumulators fit in TMEM; CTA pairs / weights in TMEM / prefetch measured, no gain: DESIGN.md 4.1); "
                            "`achieved` counts its GEMM's flops, hbm_* its own algorithmic bytes (3 tensor passes + z/dz)")
        roofline["hbm_GBps"], roofline["hbm_frac"] = kernels[top]["hbm_GBps"], kernels[top]["hbm_frac"]

    cpu = None
    if not a.no_cpu_baseline:
        rate, cores, sec, kind = cpu_reference_rate(a, reps=3)
        what = CPU_WHAT[kind] if a.stack == "film" else "DGL-formula GCN2Conv stack on CPU (oracle/dgl_shim)"
        cpu = {"value": rate, "unit": "graphs/s", "cores": cores, "kind": kind,
               "sample": f"{what}, fp32, fwd+bwd of {a.cpu_graphs} graphs, best of 3 after 1 warm-up ({a.cpu_graphs / rate * 1e3:.1f} ms)"}

    line = {"metric": METRIC, "value": value, "unit": "graphs/s", "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": a.dtype, "data": "synthetic",
            "config": config_of(a, world), "numa_node_rank0": numa,
            "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clk}
    if comm is not None:
        line["comm"] = comm
        line["grad_sync_check"] = comm["grad_sync_check"]
    if other is not None:
        line["other_configs"] = other
    if torch_gpu is not None:
        line["torch_gpu_baseline"] = torch_gpu
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
