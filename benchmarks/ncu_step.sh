#!/bin/bash
# ncu evidence for one timed step of the bench workload, as profiles/README.md describes:
#   1. plain run (must exit 0), 2. launch list (gpu__time_duration.sum of every launch), 3. --set full of the launches of the LAST step
#      (located in the launch list by the step's first kernel, k_mlp1_fwd: two per step), exported as text + raw CSV.
# Usage: benchmarks/ncu_step.sh <tag>     (writes gpurun_out/<tag>_*)
set -u
TAG=${1:-step}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-other-configs"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches.log 2>&1
read SKIP COUNT < <(python - <<PY
import csv
rows = [r for r in csv.reader(open("gpurun_out/${TAG}_launches.csv")) if len(r) > 5]
hdr = rows[0]; k = hdr.index("Kernel Name")
names = [r[k] for r in rows[1:] if r[0].isdigit()]
first = [i for i, n in enumerate(names) if "k_mlp1_fwd" in n]
print(first[-2], len(names) - first[-2])
PY
)
echo "launches of the last step: skip $SKIP count $COUNT"
ncu --set full --import-source on --clock-control none -s $SKIP -c $COUNT -o /tmp/${TAG}_full $CMD > gpurun_out/${TAG}_full.log 2>&1
ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_full.ncu-rep --page details > gpurun_out/${TAG}_full_details.txt 2>/dev/null
tail -2 gpurun_out/${TAG}_full.log
