#!/bin/bash
# round 2, GPU call M (2 GPUs): route modes -- tests, N-rank identity, bench peer (copy engines) / peer-kernel / nccl
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/m_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/m_pytest.log; grep -E "^(E  |FAILED)" gpurun_out/m_pytest.log | head
timeout 900 $TR 29511 tools/multirank_check.py --samples 10007 --tracks 40 > gpurun_out/m_multirank_${N}.json 2> gpurun_out/m_multirank_${N}.err
tail -1 gpurun_out/m_multirank_${N}.json | cut -c1-1000
for g in peer peer-kernel nccl; do
  extra="--no-e2e"; [ $g = peer ] && extra=""
  timeout 900 $TR 29512 bench.py --gpus $N --steps 8 --warmup 3 --gather $g $extra > gpurun_out/m_bench_${N}gpu_$g.json 2> gpurun_out/m_bench_${N}gpu_$g.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/m_bench_${N}gpu_$g.json"))
    print("N=%d $g: value %.0f e2e %s ms/step %.2f parity %s gather %s" % (d["n_gpus"], d["value"], d["e2e"] and round(d["e2e"]["value"]), d["ms_per_step"], d["parity_check"], d["gather_check"]))
except Exception as e:
    print("bench $g failed", e); print(open("gpurun_out/m_bench_${N}gpu_$g.err").read()[-1200:])
PY
done
