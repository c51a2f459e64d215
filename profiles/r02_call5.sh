#!/bin/bash
# Round 2, GPU call 5 (N B200s, N = $1): the driver-shaped scaling point at N GPUs -- weak scaling
# (with the N-rank == 1-process parity check, the developed regime and, N = 4, the rings
# sub-record = BASELINE configs[4]; N = 8, configs[2] as quoted) and the strong-scaling point
# of the 10 M case.
set -u
N=${1:-8}
mkdir -p gpurun_out
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513"
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2])); r = d['roofline']
    print('%s ms/step %.4f value %.4g particles %d launches/step %.1f e2e %s' % (sys.argv[1], d['ms_per_step'], d['value'], d['config']['particles'], d['launches_per_step'], d['e2e'].get('ms_per_step')))
    print('   halo', d['config']['halo'])
    p = d['config'].get('multi_gpu_parity')
    if p: print('   parity', p['ok'], max(p['max_scaled_error'].values()), p['halo'])
    print('   per_rank', d['config']['per_rank'])
    if d.get('developed'): print('   developed', d['developed'])
    if d.get('configs2_as_quoted'): print('   configs2_as_quoted', d['configs2_as_quoted'])
    for k, v in (d.get('extra') or {}).items():
        print('   extra', k, 'ms/step %.3f value %.4g particles %s' % (v['ms_per_step'], v['value'], v['config'].get('particles')), v.get('halo'))
except Exception as e:
    print(sys.argv[1], 'failed', e)
PY
}
echo "== N=$N weak (driver shape: --steps 20 --warmup 5)"
timeout 900 $T bench.py --gpus $N --steps 20 --warmup 5 > $O/r02f_n${N}_weak.json 2> $O/r02f_n${N}_weak.err
show "N=$N weak" $O/r02f_n${N}_weak.json
grep -v "^\*\*\*\|OMP_NUM_THREADS\|^NCCL version\|^$" $O/r02f_n${N}_weak.err | tail -5
echo "== N=$N weak, 100 steps after 50 (BASELINE protocol)"
timeout 900 $T bench.py --gpus $N --steps 100 --warmup 50 --no-parity --no-developed --no-extras --e2e-steps 3 > $O/r02f_n${N}_weak100.json 2> $O/r02f_n${N}_weak100.err
show "N=$N weak100" $O/r02f_n${N}_weak100.json
echo "== N=$N strong (10 M case)"
timeout 900 $T bench.py --gpus $N --scaling strong --steps 20 --warmup 5 --e2e-steps 2 --no-developed --no-parity > $O/r02f_n${N}_strong.json 2> $O/r02f_n${N}_strong.err
show "N=$N strong" $O/r02f_n${N}_strong.json
grep -v "^\*\*\*\|OMP_NUM_THREADS\|^NCCL version\|^$" $O/r02f_n${N}_strong.err | tail -5
