#!/bin/bash
# round 2, GPU call Z (1 GPU): ncu launch list of the bench command on the final tree
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/z_launches.csv python bench.py --steps 2 --warmup 1 --batches-per-step 4 --no-cpu-baseline --no-e2e > gpurun_out/z_bench.log 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(l for l in open("gpurun_out/z_launches.csv") if l.startswith('"'))]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
t=collections.defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    k=r[ki].split("(")[0]; t[k][0]+=1; t[k][1]+=float(r[vi].replace(",",""))
for k,(n,ns) in sorted(t.items(), key=lambda x:-x[1][1])[:14]: print("%-60s n=%4d total %.3f ms avg %.4f ms" % (k[:60], n, ns/1e6, ns/1e6/n))
PY
