#!/bin/bash
# round 2, GPU call D (8 GPUs): the north-star run as stated (1e6 samples x 10k segments x 1000 tracks from BED
# files through the CLI on 8 GPUs), BASELINE configs 4 and 5 on 8 / 4 GPUs, bench --gpus 8, N-rank identity at 8
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node"
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > gpurun_out/d_gpus.txt
python tools/make_bed.py /tmp/ns 2> gpurun_out/d_make_bed.err
ARGS="--segments=/tmp/ns/segments.bed --annotations=/tmp/ns/annotations.bed --workspace=/tmp/ns/workspace.bed --ignore-segment-tracks --counter=nucleotide-overlap --random-seed=1 --qvalue-method=BH"
timeout 900 python tools/run_cli_timed.py --gpus $N --label ns_1e6_${N}gpu -- $ARGS --num-samples=1000000 > gpurun_out/d_cli_ns_1e6_${N}gpu.json
cut -c1-700 gpurun_out/d_cli_ns_1e6_${N}gpu.json
timeout 900 python tools/run_cli_timed.py --gpus 1 --label ns_1e6_1gpu -- $ARGS --num-samples=1000000 > gpurun_out/d_cli_ns_1e6_1gpu.json
cut -c1-700 gpurun_out/d_cli_ns_1e6_1gpu.json
timeout 900 $TR $N --master-addr 127.0.0.1 --master-port 29541 tools/baseline_configs.py c4full c5 > gpurun_out/d_configs_${N}gpu.json 2> gpurun_out/d_configs_${N}gpu.err
cat gpurun_out/d_configs_${N}gpu.json | cut -c1-600
timeout 900 $TR 4 --master-addr 127.0.0.1 --master-port 29542 tools/baseline_configs.py c5 > gpurun_out/d_configs_4gpu.json 2> gpurun_out/d_configs_4gpu.err
cat gpurun_out/d_configs_4gpu.json | cut -c1-600
timeout 900 $TR $N --master-addr 127.0.0.1 --master-port 29543 tools/multirank_check.py --samples 10007 --tracks 40 > gpurun_out/d_multirank_${N}.json 2> gpurun_out/d_multirank_${N}.err
tail -1 gpurun_out/d_multirank_${N}.json
timeout 900 $TR $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/d_bench_${N}gpu.json 2> gpurun_out/d_bench_${N}gpu.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/d_bench_${N}gpu.json"))
    print("N=%d value %.0f e2e %.0f ms/step %.2f parity %s gather %s" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"], d["parity_check"], d["gather_check"]))
except Exception as e:
    print("bench failed", e)
PY
