#!/bin/bash
# round 2, GPU call J (1 GPU): compute-sanitizer on the kernels smoke() does not reach (K2 contig merge, K5 statistics,
# shift kernel, device preparation, output routes), after the full test suite and a bench line
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/j_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/j_pytest.log; grep -E "^(E  |FAILED)" gpurun_out/j_pytest.log | head -20
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/j_bench_ns.json 2> gpurun_out/j_bench_ns.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/j_bench_ns.json")); r=d["roofline"]
    print("ns: value %.0f e2e %.0f ms/step %.2f count %.3f place %.3f parity %s l2frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r["kernel_ms"], r["other_kernels"]["place_kernel_ms"], d["parity_check"], r["frac"]))
except Exception as e:
    print("failed", e); print(open("gpurun_out/j_bench_ns.err").read()[-800:])
PY
SEL='tests/test_gpu_parity.py::test_run_matches_oracle tests/test_gpu_parity.py::test_column_stats_match_oracle tests/test_gpu_parity.py::test_place_problem_matches_oracle tests/test_gpu_prep.py::test_lists_restrict_collapse_select tests/test_gpu_round2.py::test_output_routes_deliver_rows_and_column_blocks tests/test_gpu_round2.py::test_overflow_growth_is_exercised'
timeout 1500 compute-sanitizer --tool memcheck --log-file gpurun_out/j_memcheck.log python -m pytest $SEL -x -q -p no:cacheprovider > gpurun_out/j_memcheck.out 2>&1
tail -2 gpurun_out/j_memcheck.log; tail -2 gpurun_out/j_memcheck.out
timeout 1500 compute-sanitizer --tool racecheck --log-file gpurun_out/j_racecheck.log python -m pytest tests/test_gpu_parity.py::test_run_matches_oracle tests/test_gpu_parity.py::test_column_stats_match_oracle tests/test_gpu_round2.py::test_output_routes_deliver_rows_and_column_blocks -x -q -p no:cacheprovider > gpurun_out/j_racecheck.out 2>&1
tail -2 gpurun_out/j_racecheck.log; tail -2 gpurun_out/j_racecheck.out
BENCH="python bench.py --steps 1 --warmup 1 --batches-per-step 2 --no-cpu-baseline --no-e2e --no-checks"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:count_kernel -s 2 -c 1 -f -o gpurun_out/j_count $BENCH > gpurun_out/j_ncu_count.log 2>&1
