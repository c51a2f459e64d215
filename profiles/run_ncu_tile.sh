#!/bin/bash
TAG=${1:-r01b}
mkdir -p gpurun_out
B200SPH_DEBUG=1 B200SPH_PAIR_KERNEL=tile ncu --set full --clock-control none --import-source on -k regex:k_pair_tile -s 6 -c 1 \
    -f -o gpurun_out/pair_${TAG} \
    python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ncu_full_${TAG}.log 2>&1
grep b200sph gpurun_out/ncu_full_${TAG}.log | head -3
