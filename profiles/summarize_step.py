"""Turns what benchmarks/ncu_step.sh <tag> brought back in gpurun_out/ into the committed summaries:
   profiles/<out>_launches.txt, <out>_ncu_full.txt, <out>_ncu_dx_film_bwd_details.txt and the per-family DRAM bytes for traffic_r02.json.
   Usage: python profiles/summarize_step.py r2v r02v"""
import csv
import json
import re
import sys

tag, out = sys.argv[1], sys.argv[2]
G = "gpurun_out/"

# ---- launch list of the last step
rows = [r for r in csv.reader(open(f"{G}{tag}_launches.csv")) if len(r) > 5]
hdr = rows[0]
k, v, g, b = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
data = [r for r in rows[1:] if r[0].isdigit()]
first = [i for i, r in enumerate(data) if "k_mlp1_fwd" in r[k]]
last = data[first[-2]:]
tot = sum(float(r[v]) for r in last)
short = lambda n: re.sub(r"\(.*", "", n).replace("void ", "")
lines = [f"ncu launch list of the LAST timed step (benchmarks/ncu_step.sh {tag}: `ncu --metrics gpu__time_duration.sum --clock-control none --csv",
         f"python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-other-configs`, B200; {len(data)} launches in the process, the step's {len(last)}).",
         "Cold caches, serialised: the SHARES must agree with bench.py's `kernels` block, not the absolute times.", ""]
fam = {}
for i, r in enumerate(last):
    t = float(r[v])
    lines.append(f"{i:2d}  {short(r[k]):64s} grid {r[g]:>14s} block {r[b]:>12s} {t / 1e3:9.1f} us  {100 * t / tot:5.1f} %")
    fam[short(r[k])] = fam.get(short(r[k]), 0) + t
lines += [f"sum {tot / 1e3:.1f} us", "",
          "shares by kernel: " + ", ".join(f"{n.split('::')[-1]} {100 * t / tot:.1f} %" for n, t in sorted(fam.items(), key=lambda kv: -kv[1])[:6])]
open(f"profiles/{out}_launches.txt", "w").write("\n".join(lines) + "\n")

# ---- ncu --set full of the same launches
rows = list(csv.reader(open(f"{G}{tag}_full_raw.csv")))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
num = lambda r, key: float(r[ix[key]].replace(",", "")) if key in ix and r[ix[key]] not in ("", "n/a") else float("nan")
nbytes = lambda r, key: num(r, key) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(units[ix[key]], 1)
usec = lambda r, key: num(r, key) * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(units[ix[key]], 1)
tab, byt = [], {}
for i, r in enumerate(data):
    name = short(r[ix["Kernel Name"]])[:60]
    t, rd, wr = usec(r, "gpu__time_duration.sum"), nbytes(r, "dram__bytes_read.sum"), nbytes(r, "dram__bytes_write.sum")
    tp = num(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    tab.append(f"{i:2d} {name:60s} {t:9.1f} us  dram {rd / 1e9:6.3f}+{wr / 1e9:6.3f} GB  tensor-pipe {tp:5.1f}%  grid {r[ix['launch__grid_size']]:>6s} x "
               f"{r[ix['launch__block_size']]:>4s}  regs {num(r, 'launch__registers_per_thread'):.0f}")
    byt.setdefault(i, rd + wr)
open(f"profiles/{out}_ncu_full.txt", "w").write(
    f"ncu --set full --import-source on --clock-control none of the same {len(data)} launches (benchmarks/ncu_step.sh {tag}; bench workload G=4096, K=5, F=512,\n"
    "bf16, two layers + compress).  Columns: duration, DRAM bytes read + written, tensor-pipe activity, grid x block, registers.\n"
    + "\n".join(tab) + f"\nsum of launch durations (ncu, cold caches, serialised): {sum(usec(r, 'gpu__time_duration.sum') for r in data):.1f} us\n")
det = open(f"{G}{tag}_full_details.txt").read()
i = det.find("k_dx_film_bwd_res")
blk = det[det.rfind("\n", 0, i) + 1:]
nxt = blk.find("Context 1, Stream", blk.find("Context 1, Stream") + 50)
open(f"profiles/{out}_ncu_dx_film_bwd_details.txt", "w").write(blk[:blk.rfind("\n", 0, nxt)] if nxt > 0 else blk[:20000])
print("\n".join(tab))
print(json.dumps({i: round(x) for i, x in byt.items()}))
