#!/bin/bash
# round 2, GPU call B: parity of the S/C grid index + bin width sweep
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/b_pytest.log 2>&1
tail -5 gpurun_out/b_pytest.log
for sh in 8 9 10 11; do
  GATB_BIN_SHIFT=$sh timeout 600 python bench.py --steps 6 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/b_bench_shift$sh.json 2> gpurun_out/b_bench_shift$sh.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/b_bench_shift$sh.json")); r=d["roofline"]
    print("shift $sh: value %.0f  count %.3f ms place %.3f ms  tested/seg %.2f pairs/seg %.2f index %.0f MB parity %s l2frac %.3f" % (d["value"], r["kernel_ms"], r["other_kernels"]["place_kernel_ms"], r["work"]["entries_tested_per_segment"], r["work"]["pairs_per_segment"], r["work"]["index_bytes"]/1e6, d["parity_check"], r["frac"]))
except Exception as e:
    print("shift $sh failed", e); print(open("gpurun_out/b_bench_shift$sh.err").read()[-1500:])
PY
done
