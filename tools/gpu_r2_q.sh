#!/bin/bash
# round 2, GPU call Q (2 GPUs): N-rank identity of gat_b200.run with the streamed statistics (all-gather and column
# layouts, both transports), bench --gpus 2
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/multirank_check.py --samples 10007 --tracks 40 > gpurun_out/q_multirank_${N}.json 2> gpurun_out/q_multirank_${N}.err
tail -1 gpurun_out/q_multirank_${N}.json; tail -3 gpurun_out/q_multirank_${N}.err
timeout 600 $TR tools/multirank_check.py --samples 3001 --tracks 20 --isochores > gpurun_out/q_multirank_iso_${N}.json 2> gpurun_out/q_multirank_iso_${N}.err
tail -1 gpurun_out/q_multirank_iso_${N}.json; tail -3 gpurun_out/q_multirank_iso_${N}.err
timeout 600 $TR bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/q_bench_${N}gpu.json 2> gpurun_out/q_bench_${N}gpu.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/q_bench_${N}gpu.json"))
    print("N=%d value %.0f e2e %.0f ms/step %.2f parity %s gather %s stats %s" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"], d["parity_check"], d["gather_check"], d["roofline"].get("statistics_kernel", {}).get("frac")))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/q_bench_${N}gpu.err").read()[-1500:])
PY
