#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
N=${1:-2}
GATB_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 tools/baseline_configs.py c4full c4full c5 > gpurun_out/y_configs_${N}gpu.json 2> gpurun_out/y_configs_${N}gpu.err
python - <<PY
import json
for line in open("gpurun_out/y_configs_${N}gpu.json"):
    if line.startswith("{"):
        d = json.loads(line); print(d["config"], "gpus", d["gpus"], "prep_s", d["prep_s"], "run_s", d["run_s"], "md5", d["table_md5"])
PY
grep -i "timing\|observed\|sampling" gpurun_out/y_configs_${N}gpu.err | head -8
