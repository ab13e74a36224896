#!/bin/bash
# round 2, GPU call I (1 GPU): parity + bench after the LONG template; ncu capture of the final count kernel
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/i_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/i_pytest.log; grep -E "^(E  |FAILED)" gpurun_out/i_pytest.log | head -20
timeout 900 python tools/stress_count.py > gpurun_out/i_stress_count.txt 2>&1; tail -2 gpurun_out/i_stress_count.txt
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/i_bench_ns.json 2> gpurun_out/i_bench_ns.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/i_bench_ns.json")); r=d["roofline"]
    print("ns: value %.0f e2e %.0f ms/step %.2f count %.3f place %.3f parity %s l2frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r["kernel_ms"], r["other_kernels"]["place_kernel_ms"], d["parity_check"], r["frac"]))
except Exception as e:
    print("failed", e); print(open("gpurun_out/i_bench_ns.err").read()[-800:])
PY
BENCH="python bench.py --steps 1 --warmup 1 --batches-per-step 2 --no-cpu-baseline --no-e2e --no-checks"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/i_launches.csv $BENCH > gpurun_out/i_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:count_kernel -s 2 -c 1 -f -o gpurun_out/i_count $BENCH > gpurun_out/i_ncu_count.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:place_kernel -s 2 -c 1 -f -o gpurun_out/i_place $BENCH > gpurun_out/i_ncu_place.log 2>&1
