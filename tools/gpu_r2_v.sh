#!/bin/bash
# round 2, GPU call V (1 GPU): 2-D launch grid of the placement kernels -- tests, ns and c3 bench lines
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/v_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/v_pytest.log; grep -E "^(E  |FAILED)" gpurun_out/v_pytest.log | head -20
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/v_bench_ns.json 2> gpurun_out/v_bench_ns.err
timeout 300 python bench.py --config c3 --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/v_bench_c3.json 2> gpurun_out/v_bench_c3.err
python - <<PY
import json
for f in ("gpurun_out/v_bench_ns.json", "gpurun_out/v_bench_c3.json"):
    try:
        d=json.load(open(f)); r=d["roofline"]
        print(f, "value %.0f e2e %s ms/step %.2f count %.3f place %.3f merge %.3f parity %s" % (d["value"], d["e2e"] and round(d["e2e"]["value"]), d["ms_per_step"], r["kernel_ms"], r["other_kernels"]["place_kernel_ms"], r["other_kernels"]["contig_merge_kernel_ms"], d["parity_check"]))
    except Exception as e:
        print("failed", f, e); print(open(f.replace(".json", ".err")).read()[-600:])
PY
