#!/bin/bash
# round 2, GPU call M (1 GPU): L2 discard of the placement kernels' dead scratch -- GPU tests, bench lines with the
# discard on / off (north-star shape and the isochore configuration), DRAM bytes of the placement kernels under ncu
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/m_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/m_pytest.log; grep -E "^(E  |FAILED)" gpurun_out/m_pytest.log | head -20
line() {
python - "$1" <<PY
import json, sys
try:
    d=json.load(open(sys.argv[1])); r=d["roofline"]
    print(sys.argv[1], "value %.0f e2e %s ms/step %.2f count %.3f place %.3f merge %.3f parity %s" % (d["value"], d["e2e"] and round(d["e2e"]["value"]), d["ms_per_step"], r["kernel_ms"], r["other_kernels"]["place_kernel_ms"], r["other_kernels"]["contig_merge_kernel_ms"], d["parity_check"]))
except Exception as e:
    print("failed", sys.argv[1], e)
PY
}
for disc in 1 0; do
  GATB_DISCARD=$disc timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/m_bench_ns_d$disc.json 2> gpurun_out/m_bench_ns_d$disc.err
  line gpurun_out/m_bench_ns_d$disc.json
  GATB_DISCARD=$disc timeout 300 python bench.py --config c3 --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/m_bench_c3_d$disc.json 2> gpurun_out/m_bench_c3_d$disc.err
  line gpurun_out/m_bench_c3_d$disc.json
done
M=gpu__time_duration.sum,dram__bytes_write.sum,dram__bytes_read.sum,smsp__inst_executed.sum
for disc in 1 0; do
  GATB_DISCARD=$disc timeout 300 ncu --metrics $M --clock-control none -k regex:"place_kernel|contig_merge" -c 4 --csv --log-file gpurun_out/m_ncu_ns_d$disc.csv \
     python bench.py --steps 1 --warmup 1 --batches-per-step 2 --no-cpu-baseline --no-e2e --no-checks > /dev/null 2>&1
  GATB_DISCARD=$disc timeout 300 ncu --metrics $M --clock-control none -k regex:"place_kernel|contig_merge" -c 6 --csv --log-file gpurun_out/m_ncu_c3_d$disc.csv \
     python bench.py --config c3 --steps 1 --warmup 1 --batches-per-step 2 --no-cpu-baseline --no-e2e --no-checks > /dev/null 2>&1
done
python - <<PY
import csv, glob
for f in sorted(glob.glob("gpurun_out/m_ncu_*.csv")):
    rows = [r for r in csv.reader(l for l in open(f) if l.startswith('"'))]
    if not rows: print(f, "empty"); continue
    h = rows[0]; ki, mi, vi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
    idi = h.index("ID")
    agg = {}
    for r in rows[1:]:
        agg.setdefault((r[idi], r[ki][:24]), {})[r[mi]] = r[vi]
    for k, v in list(agg.items())[-3:]:
        print(f, k, v)
PY
