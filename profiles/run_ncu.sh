#!/bin/bash
# Run on the GPU box (under gpurun): launch list + full capture of the pair kernel.
# usage: bash profiles/run_ncu.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv \
    --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ncu_bench_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pair -s 6 -c 2 \
    -f -o gpurun_out/pair_${TAG} \
    python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out/
