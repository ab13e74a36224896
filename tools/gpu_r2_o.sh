#!/bin/bash
# round 2, GPU call O (1 GPU): streamed statistics after the per-chunk overhead cut -- K5 benchmark, ncu summary of both
# kernels, and the north-star CLI run on one GPU (statistics phase, table checksum)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 120 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q -k "stats" ) > gpurun_out/o_pytest_stats.log 2>&1
grep -E "passed|failed" gpurun_out/o_pytest_stats.log
if ! grep -q passed gpurun_out/o_pytest_stats.log || grep -q failed gpurun_out/o_pytest_stats.log; then tail -30 gpurun_out/o_pytest_stats.log; exit 1; fi
timeout 300 python tools/stats_bench.py > gpurun_out/o_stats_bench_1e6x1000.json 2> gpurun_out/o_stats_bench.err; cat gpurun_out/o_stats_bench_1e6x1000.json
timeout 300 python tools/stats_bench.py --cols 125 > gpurun_out/o_stats_bench_1e6x125.json 2>> gpurun_out/o_stats_bench.err; cat gpurun_out/o_stats_bench_1e6x125.json
timeout 300 python tools/stats_bench.py --cols 50 --samples 4000000 > gpurun_out/o_stats_bench_4e6x50.json 2>> gpurun_out/o_stats_bench.err; cat gpurun_out/o_stats_bench_4e6x50.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stats_stream_kernel -c 2 -o gpurun_out/o_stats_stream -f python tools/stats_bench.py --samples 400000 --reps 1 > /dev/null 2> gpurun_out/o_ncu.err
python tools/make_bed.py /tmp/ns 2> gpurun_out/o_make_bed.err
ARGS="--segments=/tmp/ns/segments.bed --annotations=/tmp/ns/annotations.bed --workspace=/tmp/ns/workspace.bed --ignore-segment-tracks --counter=nucleotide-overlap --random-seed=1 --qvalue-method=BH --num-samples=1000000"
timeout 600 python tools/run_cli_timed.py --gpus 1 --label ns_1e6_1gpu -- $ARGS > gpurun_out/o_cli_ns_1e6_1gpu.json
cut -c1-700 gpurun_out/o_cli_ns_1e6_1gpu.json
