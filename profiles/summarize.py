#!/usr/bin/env python
"""Turn ncu outputs (brought back in gpurun_out/) into the small text summaries committed under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_X.csv  > profiles/rNN_launches.txt
    python profiles/summarize.py full     gpurun_out/prof_X.ncu-rep  > profiles/rNN_ncu_full.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [re.sub(r"\(.*", "", r["Kernel Name"])[:70] for r in rows]
    dur = [float(r["Metric Value"].replace(",", "")) for r in rows]
    start = 0
    for i, n in enumerate(names):
        if "k_build_csr" in n:
            start = i
    tot = collections.OrderedDict()
    for n, d in zip(names[start + 1:], dur[start + 1:]):
        tot.setdefault(n, [0, 0.0])
        tot[n][0] += 1
        tot[n][1] += d
    total = sum(v[1] for v in tot.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none : {len(rows)} launches captured; "
          f"{len(names) - start - 1} after the CSR build (warm-up + timed steps of bench.py)")
    print("# per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes")
    print(f"{'kernel':72s} {'launches':>8s} {'total ms':>10s} {'share':>7s}")
    for n, (c, d) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{n:72s} {c:8d} {d / 1e6:10.3f} {100 * d / total:6.1f}%")
    print(f"{'TOTAL':72s} {sum(v[0] for v in tot.values()):8d} {total / 1e6:10.3f}")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
    kn = hdr.index("Kernel Name")
    print(f"# ncu --set full --clock-control none ({path}); one block per captured launch")
    for r in data:
        print(f"\n== {r[kn][:110]}")
        for w, i in idx:
            print(f"   {w:72s} {r[i]:>16s} {units[i]}")
        try:
            t = float(r[hdr.index('gpu__time_duration.sum')])
            tu = units[hdr.index('gpu__time_duration.sum')]
            rd, wr = float(r[hdr.index('dram__bytes_read.sum')]), float(r[hdr.index('dram__bytes_write.sum')])
            ru = units[hdr.index('dram__bytes_read.sum')]
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[ru]
            wu = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[hdr.index('dram__bytes_write.sum')]]
            ts = {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}[tu]
            print(f"   {'-> DRAM traffic (read+write) / duration':72s} {(rd * scale + wr * wu) / (t * ts) / 1e9:16.1f} GB/s   "
                  f"traffic {(rd * scale + wr * wu) / 1e9:.3f} GB")
        except Exception:
            pass


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
