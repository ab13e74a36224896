#!/bin/bash
# round 2, GPU call A: full GPU test suite, a bench line, compute-sanitizer on the smoke problem, ncu baseline captures
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/a_gpu.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/a_pytest.log 2>&1
tail -5 gpurun_out/a_pytest.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_ns.json 2> gpurun_out/a_bench_ns.err
cat gpurun_out/a_bench_ns.json | cut -c1-1500
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/a_memcheck.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_memcheck.out 2>&1
tail -3 gpurun_out/a_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/a_racecheck.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_racecheck.out 2>&1
tail -3 gpurun_out/a_racecheck.log
BENCH="python bench.py --steps 1 --warmup 1 --batches-per-step 2 --no-cpu-baseline --no-e2e --no-checks"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/a_launches.csv $BENCH > gpurun_out/a_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:count_kernel -s 2 -c 1 -f -o gpurun_out/a_count $BENCH > gpurun_out/a_ncu_count.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:place_kernel -s 2 -c 1 -f -o gpurun_out/a_place $BENCH > gpurun_out/a_ncu_place.log 2>&1
ls -la gpurun_out
