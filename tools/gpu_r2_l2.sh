#!/bin/bash
# round 2, GPU call L2 (8 GPUs): final numbers with the copy-engine routes -- north-star CLI run, bench --gpus 8 (peer / nccl), configs 4 / 5
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node"
python tools/make_bed.py /tmp/ns 2> gpurun_out/l2_make_bed.err
ARGS="--segments=/tmp/ns/segments.bed --annotations=/tmp/ns/annotations.bed --workspace=/tmp/ns/workspace.bed --ignore-segment-tracks --counter=nucleotide-overlap --random-seed=1 --qvalue-method=BH --num-samples=1000000"
timeout 900 python tools/run_cli_timed.py --gpus $N --label ns_1e6_${N}gpu_peer -- $ARGS > gpurun_out/l2_cli_ns_1e6_${N}gpu_peer.json
cut -c1-620 gpurun_out/l2_cli_ns_1e6_${N}gpu_peer.json
timeout 900 $TR $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/l2_bench_${N}gpu_peer.json 2> gpurun_out/l2_bench_${N}gpu_peer.err
timeout 900 $TR $N --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus $N --steps 10 --warmup 3 --gather nccl --no-e2e > gpurun_out/l2_bench_${N}gpu_nccl.json 2> gpurun_out/l2_bench_${N}gpu_nccl.err
python - <<PY
import json
for tag in ("peer", "nccl"):
    try:
        d=json.load(open("gpurun_out/l2_bench_${N}gpu_%s.json" % tag))
        print("N=%d %s: value %.0f e2e %s ms/step %.2f parity %s gather %s" % (d["n_gpus"], tag, d["value"], d["e2e"] and round(d["e2e"]["value"]), d["ms_per_step"], d["parity_check"], d["gather_check"]))
    except Exception as e:
        print("bench", tag, "failed", e); print(open("gpurun_out/l2_bench_${N}gpu_%s.err" % tag).read()[-1200:])
PY
timeout 900 $TR $N --master-addr 127.0.0.1 --master-port 29541 tools/baseline_configs.py c4full c5 > gpurun_out/l2_configs_${N}gpu.json 2> gpurun_out/l2_configs_${N}gpu.err
cut -c1-330 gpurun_out/l2_configs_${N}gpu.json
