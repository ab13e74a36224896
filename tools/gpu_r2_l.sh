#!/bin/bash
# round 2, GPU call L (8 GPUs): final multi-GPU evidence -- the north-star run as stated with the peer transport and
# device-side preparation (and with NCCL for comparison), bench --gpus 8 with both gathers, configs 4 / 5, N-rank identity
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node"
python tools/make_bed.py /tmp/ns 2> gpurun_out/l_make_bed.err
ARGS="--segments=/tmp/ns/segments.bed --annotations=/tmp/ns/annotations.bed --workspace=/tmp/ns/workspace.bed --ignore-segment-tracks --counter=nucleotide-overlap --random-seed=1 --qvalue-method=BH --num-samples=1000000"
timeout 900 python tools/run_cli_timed.py --gpus $N --label ns_1e6_${N}gpu_peer -- $ARGS > gpurun_out/l_cli_ns_1e6_${N}gpu_peer.json
cut -c1-600 gpurun_out/l_cli_ns_1e6_${N}gpu_peer.json
GATB_TRANSPORT=nccl timeout 900 python tools/run_cli_timed.py --gpus $N --label ns_1e6_${N}gpu_nccl --port 29532 -- $ARGS > gpurun_out/l_cli_ns_1e6_${N}gpu_nccl.json
cut -c1-600 gpurun_out/l_cli_ns_1e6_${N}gpu_nccl.json
timeout 900 python tools/run_cli_timed.py --gpus 1 --label ns_1e6_1gpu -- $ARGS > gpurun_out/l_cli_ns_1e6_1gpu.json
cut -c1-600 gpurun_out/l_cli_ns_1e6_1gpu.json
timeout 900 $TR $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/l_bench_${N}gpu_peer.json 2> gpurun_out/l_bench_${N}gpu_peer.err
timeout 900 $TR $N --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus $N --steps 10 --warmup 3 --gather nccl --no-e2e > gpurun_out/l_bench_${N}gpu_nccl.json 2> gpurun_out/l_bench_${N}gpu_nccl.err
python - <<PY
import json
for tag in ("peer", "nccl"):
    try:
        d=json.load(open("gpurun_out/l_bench_${N}gpu_%s.json" % tag))
        print("N=%d %s: value %.0f e2e %s ms/step %.2f parity %s gather %s" % (d["n_gpus"], tag, d["value"], d["e2e"] and round(d["e2e"]["value"]), d["ms_per_step"], d["parity_check"], d["gather_check"]))
    except Exception as e:
        print("bench", tag, "failed", e); print(open("gpurun_out/l_bench_${N}gpu_%s.err" % tag).read()[-1200:])
PY
timeout 900 $TR $N --master-addr 127.0.0.1 --master-port 29541 tools/baseline_configs.py c4full c5 > gpurun_out/l_configs_${N}gpu.json 2> gpurun_out/l_configs_${N}gpu.err
cut -c1-400 gpurun_out/l_configs_${N}gpu.json
timeout 900 $TR $N --master-addr 127.0.0.1 --master-port 29543 tools/multirank_check.py --samples 10007 --tracks 40 > gpurun_out/l_multirank_${N}.json 2> gpurun_out/l_multirank_${N}.err
tail -1 gpurun_out/l_multirank_${N}.json | cut -c1-900
