#!/bin/bash
# round 2, GPU call K (1 GPU): quick check -- full GPU tests + one bench line (flags after --)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/k_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/k_pytest.log; grep -E "^(E  |FAILED)" gpurun_out/k_pytest.log | head -20
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/k_bench.json")); r=d["roofline"]
    print("value %.0f e2e %s ms/step %.2f count %.3f place %.3f merge %.3f parity %s l2frac %.3f" % (d["value"], d["e2e"] and round(d["e2e"]["value"]), d["ms_per_step"], r["kernel_ms"], r["other_kernels"]["place_kernel_ms"], r["other_kernels"]["contig_merge_kernel_ms"], d["parity_check"], r["frac"]))
except Exception as e:
    print("failed", e); print(open("gpurun_out/k_bench.err").read()[-800:])
PY
