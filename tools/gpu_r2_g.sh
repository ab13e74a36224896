#!/bin/bash
# round 2, GPU call G (1 GPU): placement window + bucket tables -- parity (tests + stress), bench ns / isochores
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/g_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/g_pytest.log; grep -E "^(E  |FAILED)" gpurun_out/g_pytest.log | head -20
timeout 900 python tools/stress_place.py 300 > gpurun_out/g_stress_place.txt 2>&1; tail -1 gpurun_out/g_stress_place.txt
for cfg in "ns:" "iso:--isochores --counter segment-overlap" "c3:--config c3"; do
  name=${cfg%%:*}; flags=${cfg#*:}
  timeout 600 python bench.py $flags --steps 6 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/g_bench_$name.json 2> gpurun_out/g_bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/g_bench_$name.json")); r=d["roofline"]
    print("$name: value %.0f  count %.3f place %.3f merge %.3f parity %s" % (d["value"], r["kernel_ms"], r["other_kernels"]["place_kernel_ms"], r["other_kernels"]["contig_merge_kernel_ms"], d["parity_check"]))
except Exception as e:
    print("$name bench failed", e); print(open("gpurun_out/g_bench_$name.err").read()[-1500:])
PY
done
