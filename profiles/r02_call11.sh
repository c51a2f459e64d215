#!/bin/bash
# Round 2: interior launch behind the outgoing messages (B200SPH_PUSH_FIRST) on / off at N GPUs.
set -u
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514"
for pf in 1 0; do
B200SPH_PUSH_FIRST=$pf timeout 600 $T bench.py --gpus $N --steps 20 --warmup 5 --no-extras --no-developed --no-cpu --e2e-steps 2 $( [ $pf = 1 ] && echo --no-parity ) > $O/r02l_n${N}_pf$pf.json 2> $O/r02l_n${N}_pf$pf.err
python - $N $pf <<'PY'
import json, sys
N, pf = sys.argv[1:3]
try:
    d = json.load(open('gpurun_out/r02l_n%s_pf%s.json' % (N, pf)))
    print('N=%s push_first=%s ms/step %.4f value %.4g' % (N, pf, d['ms_per_step'], d['value']))
    p = d['config'].get('multi_gpu_parity')
    if p: print('   parity', p['ok'], max(p['max_scaled_error'].values()))
    print('   per_rank (ms_pair, ms_other, sent, reduced, chain, pair_wall)', [(r['ms_pair'], r['ms_other'], r.get('ms_halo_sent'), r.get('ms_halo_reduced'), r.get('ms_halo_chain'), r.get('ms_pair_wall')) for r in d['config']['per_rank']])
except Exception as e:
    print('failed', e)
PY
done
