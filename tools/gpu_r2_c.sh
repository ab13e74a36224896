#!/bin/bash
# round 2, GPU call C (2 GPUs): N-rank identity of gat_b200.run, bench --gpus 2 with gather_check, tests
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/c_pytest.log 2>&1
tail -4 gpurun_out/c_pytest.log
timeout 900 $TR tools/multirank_check.py --samples 10007 --tracks 40 > gpurun_out/c_multirank_${N}.json 2> gpurun_out/c_multirank_${N}.err
tail -1 gpurun_out/c_multirank_${N}.json; tail -3 gpurun_out/c_multirank_${N}.err
timeout 900 $TR tools/multirank_check.py --samples 3001 --tracks 20 --isochores > gpurun_out/c_multirank_iso_${N}.json 2> gpurun_out/c_multirank_iso_${N}.err
tail -1 gpurun_out/c_multirank_iso_${N}.json; tail -3 gpurun_out/c_multirank_iso_${N}.err
timeout 900 $TR bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/c_bench_${N}gpu.json 2> gpurun_out/c_bench_${N}gpu.err
cut -c1-900 gpurun_out/c_bench_${N}gpu.json; tail -3 gpurun_out/c_bench_${N}gpu.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/c_bench_${N}gpu.json"))
    print("N=%d value %.0f e2e %.0f ms/step %.2f parity %s gather %s" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"], d["parity_check"], d["gather_check"]))
except Exception as e:
    print("bench failed", e)
PY
timeout 900 $TR bench.py --gpus $N --steps 8 --warmup 3 --gather nccl --no-e2e > gpurun_out/c_bench_${N}gpu_nccl.json 2> gpurun_out/c_bench_${N}gpu_nccl.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/c_bench_${N}gpu_nccl.json"))
    print("N=%d (nccl gather) value %.0f ms/step %.2f parity %s gather %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["parity_check"], d["gather_check"]))
except Exception as e:
    print("bench nccl failed", e)
PY
