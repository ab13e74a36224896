#!/bin/bash
# round 2, GPU call N (1 GPU): TMA-streamed column statistics -- GPU tests (bounded), K5 benchmark against the
# column-tiled kernels on the north-star matrix, bench line with roofline.statistics_kernel, ncu of the new kernel
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 120 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q -k "stats" ) > gpurun_out/n_pytest_stats.log 2>&1
grep -E "passed|failed" gpurun_out/n_pytest_stats.log; grep -E "^(E  |FAILED)" gpurun_out/n_pytest_stats.log | head -20
if grep -q -E "failed|error|Timeout|Killed" gpurun_out/n_pytest_stats.log || ! grep -q passed gpurun_out/n_pytest_stats.log; then echo "stats tests did not pass: stopping"; tail -30 gpurun_out/n_pytest_stats.log; exit 1; fi
timeout 300 python tools/stats_bench.py > gpurun_out/n_stats_bench_1e6x1000.json 2> gpurun_out/n_stats_bench.err; cat gpurun_out/n_stats_bench_1e6x1000.json; tail -3 gpurun_out/n_stats_bench.err
timeout 300 python tools/stats_bench.py --cols 125 > gpurun_out/n_stats_bench_1e6x125.json 2>> gpurun_out/n_stats_bench.err; cat gpurun_out/n_stats_bench_1e6x125.json
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/n_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/n_pytest.log; grep -E "^(E  |FAILED)" gpurun_out/n_pytest.log | head -20
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/n_bench.json")); r=d["roofline"]
    print("value %.0f e2e %s ms/step %.2f count %.3f place %.3f parity %s" % (d["value"], d["e2e"] and round(d["e2e"]["value"]), d["ms_per_step"], r["kernel_ms"], r["other_kernels"]["place_kernel_ms"], d["parity_check"]))
    print("stats", json.dumps(r.get("statistics_kernel")))
except Exception as e:
    print("failed", e); print(open("gpurun_out/n_bench.err").read()[-800:])
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stats_stream_kernel -c 2 -o gpurun_out/n_stats_stream -f python tools/stats_bench.py --samples 400000 --reps 1 > /dev/null 2> gpurun_out/n_ncu.err
ls -la gpurun_out/n_stats_stream.ncu-rep
