#!/bin/bash
# round 2, GPU call T (1 GPU): final check of the tree -- smoke, full GPU tests, default bench line, reference arm
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/t_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/t_pytest.log; grep -E "^(E  |FAILED)" gpurun_out/t_pytest.log | head -20
timeout 900 python bench.py > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/t_bench.json")); r=d["roofline"]
    print("value %.0f e2e %s ms/step %.2f steps %d count %.3f place %.3f parity %s l2frac %.3f issue %.3f stats %.3f cpu %s launches %s" % (d["value"], d["e2e"] and round(d["e2e"]["value"]), d["ms_per_step"], d["steps"], r["kernel_ms"], r["other_kernels"]["place_kernel_ms"], d["parity_check"], r["frac"], r["issue"]["frac"], r["statistics_kernel"]["frac"], d["cpu_baseline"]["value"], d.get("gpu_launches")))
except Exception as e:
    print("failed", e); print(open("gpurun_out/t_bench.err").read()[-800:])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/t_bench_ref.json 2> gpurun_out/t_bench_ref.err; cut -c1-400 gpurun_out/t_bench_ref.json
