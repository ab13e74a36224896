#!/bin/bash
# round 2, GPU call S (1 GPU): contig merge with precomputed plan + histogram during the gather -- tests, c3 bench line
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/s_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/s_pytest.log; grep -E "^(E  |FAILED)" gpurun_out/s_pytest.log | head -20
timeout 300 python bench.py --config c3 --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s_bench_c3.json 2> gpurun_out/s_bench_c3.err
timeout 300 python bench.py --isochores --counter segment-overlap --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s_bench_iso1000.json 2> gpurun_out/s_bench_iso1000.err
python - <<PY
import json
for f in ("gpurun_out/s_bench_c3.json", "gpurun_out/s_bench_iso1000.json"):
    try:
        d=json.load(open(f)); r=d["roofline"]
        print(f, "value %.0f ms/step %.2f count %.3f place %.3f merge %.3f parity %s" % (d["value"], d["ms_per_step"], r["kernel_ms"], r["other_kernels"]["place_kernel_ms"], r["other_kernels"]["contig_merge_kernel_ms"], d["parity_check"]))
    except Exception as e:
        print("failed", f, e); print(open(f.replace(".json", ".err")).read()[-600:])
PY
