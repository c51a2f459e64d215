#!/bin/bash
# Run on the GPU box (under gpurun, ONE GPU): per-launch durations of a few steps of
# both bench workloads + one `--set full` capture of the dominant kernels.
# usage: bash profiles/run_ncu.sh <tag>     (outputs land in gpurun_out/)
TAG=${1:-r01h}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv \
    --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ncu_bench_${TAG}.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file gpurun_out/launches_${TAG}_tg.csv \
    python bench.py --workload taylor_green --steps 3 --warmup 3 --no-cpu --e2e-steps 1 \
    > gpurun_out/ncu_bench_${TAG}_tg.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tvf_pass -s 4 -c 2 \
    -f -o gpurun_out/tvf_${TAG} \
    python bench.py --workload taylor_green --steps 1 --warmup 3 --no-cpu --e2e-steps 1 \
    > gpurun_out/ncu_full_${TAG}_tg.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pair_list -s 6 -c 1 \
    -f -o gpurun_out/pair_${TAG} \
    python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out/*${TAG}*
