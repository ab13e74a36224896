#!/bin/bash
# round 2, GPU call X (N GPUs): BASELINE configurations 4 / 5 (and 2, 3) through gat_b200.run on the final tree; table md5 must not depend on N
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
N=${1:-1}
if [ "$N" = "1" ]; then
  timeout 600 python tools/baseline_configs.py c2 c3 c5 c4full > gpurun_out/x_configs_1gpu.json 2> gpurun_out/x_configs_1gpu.err
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 tools/baseline_configs.py c5 c4full > gpurun_out/x_configs_${N}gpu.json 2> gpurun_out/x_configs_${N}gpu.err
fi
python - <<PY
import json
for line in open("gpurun_out/x_configs_${N}gpu.json"):
    if line.startswith("{"):
        d = json.loads(line); print(d["config"], "gpus", d["gpus"], "run_s", d["run_s"], "samples/s", d["samples_per_s"], "md5", d["table_md5"])
PY
tail -2 gpurun_out/x_configs_${N}gpu.err
