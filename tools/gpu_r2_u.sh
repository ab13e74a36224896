#!/bin/bash
# round 2, GPU call U (1 GPU): per-source-line profile of the placement and contig-merge kernels on the isochore configuration
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"place_kernel" -c 1 -o gpurun_out/u_place_c3 -f python bench.py --config c3 --steps 1 --warmup 1 --batches-per-step 2 --no-cpu-baseline --no-e2e --no-checks > /dev/null 2> gpurun_out/u_ncu1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"contig_merge" -c 1 -o gpurun_out/u_merge_c3 -f python bench.py --config c3 --steps 1 --warmup 1 --batches-per-step 2 --no-cpu-baseline --no-e2e --no-checks > /dev/null 2> gpurun_out/u_ncu2.err
ls -la gpurun_out/u_*.ncu-rep
