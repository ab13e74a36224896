#!/bin/bash
# round 2, GPU call F (1 GPU): drop-in tests; isochore configuration baseline (bench + ncu of place / merge kernels)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/f_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/f_pytest.log; grep -E "^(E  |FAILED)" gpurun_out/f_pytest.log | head -20
ISO="--isochores --counter segment-overlap --no-cpu-baseline --no-e2e"
timeout 600 python bench.py $ISO --steps 5 --warmup 2 > gpurun_out/f_bench_iso.json 2> gpurun_out/f_bench_iso.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/f_bench_iso.json")); r=d["roofline"]
    print("iso: value %.0f  count %.3f place %.3f merge %.3f parity %s" % (d["value"], r["kernel_ms"], r["other_kernels"]["place_kernel_ms"], r["other_kernels"]["contig_merge_kernel_ms"], d["parity_check"]))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/f_bench_iso.err").read()[-1500:])
PY
B="python bench.py $ISO --steps 1 --warmup 1 --batches-per-step 2 --no-checks"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:place_kernel -s 2 -c 1 -f -o gpurun_out/f_place_iso $B > gpurun_out/f_ncu_place.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:contig_merge -s 2 -c 1 -f -o gpurun_out/f_merge_iso $B > gpurun_out/f_ncu_merge.log 2>&1
ls -la gpurun_out | grep f_
