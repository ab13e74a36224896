#!/bin/bash
# round 2, GPU call E (1 GPU): device-side input preparation -- tests, and the north-star CLI run with device vs host prep
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/e_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/e_pytest.log; grep -E "^(E  |FAILED)" gpurun_out/e_pytest.log | head -20
python tools/make_bed.py /tmp/ns 2> gpurun_out/e_make_bed.err
ARGS="--segments=/tmp/ns/segments.bed --annotations=/tmp/ns/annotations.bed --workspace=/tmp/ns/workspace.bed --ignore-segment-tracks --counter=nucleotide-overlap --random-seed=1 --qvalue-method=BH --num-samples=100000"
timeout 900 python tools/run_cli_timed.py --gpus 1 --label ns_1e5_device_prep -- $ARGS > gpurun_out/e_cli_device_prep.json
cut -c1-900 gpurun_out/e_cli_device_prep.json
GATB_HOST_PREP=1 timeout 900 python tools/run_cli_timed.py --gpus 1 --label ns_1e5_host_prep -- $ARGS > gpurun_out/e_cli_host_prep.json
cut -c1-900 gpurun_out/e_cli_host_prep.json
