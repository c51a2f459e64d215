#!/bin/bash
# Run on the GPU box (under gpurun, 1 GPU): launch list of a few steps + full capture of
# the dominant kernel (k_pair_list) and of the list builder.
TAG=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ncu_bench_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pair_list -s 6 -c 1 \
    -f -o gpurun_out/pair_${TAG} \
    python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out/*${TAG}*
