#!/bin/bash
# N = 2: final numbers + host-side timeline (B200SPH_PM_PROFILE: cpu seconds per phase of the loop)
set -u
mkdir -p gpurun_out
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515"
bash profiles/r02_call8.sh 2
echo "== host timeline"
B200SPH_PM_PROFILE=1 timeout 400 $T bench.py --gpus 2 --steps 40 --warmup 10 --e2e-steps 2 --no-developed --no-parity --no-extras > $O/r02j_n2_prof.json 2> $O/r02j_n2_prof.err
grep "pm-profile" $O/r02j_n2_prof.err
python -c "
import json; d=json.load(open('$O/r02j_n2_prof.json')); print('ms/step', d['ms_per_step'])"
