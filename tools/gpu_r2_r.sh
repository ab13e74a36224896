#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 300 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q -k "stats" ) > gpurun_out/r_pytest_stats.log 2>&1
grep -E "passed|failed" gpurun_out/r_pytest_stats.log; grep -E "^(E  |FAILED)" gpurun_out/r_pytest_stats.log | head
bash tools/gpu_r2_q.sh 2
