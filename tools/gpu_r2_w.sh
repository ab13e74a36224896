#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 300 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q -k "stats" ) > gpurun_out/w_pytest_stats.log 2>&1
grep -E "passed|failed" gpurun_out/w_pytest_stats.log
timeout 300 python tools/stats_bench.py > gpurun_out/w_stats_bench_1e6x1000.json 2> gpurun_out/w_stats_bench.err; cut -c1-330 gpurun_out/w_stats_bench_1e6x1000.json
timeout 300 python tools/stats_bench.py --cols 125 > gpurun_out/w_stats_bench_1e6x125.json 2>> gpurun_out/w_stats_bench.err; cut -c1-330 gpurun_out/w_stats_bench_1e6x125.json
