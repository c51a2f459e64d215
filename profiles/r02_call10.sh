#!/bin/bash
# Round 2: the merged push / pull kernels + the agreement on its own stream, at N GPUs (N = $1):
# the multi-rank GPU tests, then the driver-shaped weak run (no extras).
set -u
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_rings_multi.py tests/test_gpu_recut.py tests/test_gpu_halo.py -m gpu -q -x > $O/r02k_n${N}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $O/r02k_n${N}_pytest.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514"
timeout 900 $T bench.py --gpus $N --steps 20 --warmup 5 --no-extras > $O/r02k_n${N}_weak.json 2> $O/r02k_n${N}_weak.err
python - $N <<'PY'
import json, sys
N = sys.argv[1]
try:
    d = json.load(open('gpurun_out/r02k_n%s_weak.json' % N))
    print('N=%s weak ms/step %.4f value %.4g particles %d launches/step %.1f e2e %s' % (N, d['ms_per_step'], d['value'], d['config']['particles'], d['launches_per_step'], d['e2e'].get('ms_per_step')))
    p = d['config'].get('multi_gpu_parity')
    if p: print('   parity', p['ok'], max(p['max_scaled_error'].values()))
    print('   per_rank (ms_pair, ms_other, sent, reduced, chain, pair_wall)', [(r['ms_pair'], r['ms_other'], r.get('ms_halo_sent'), r.get('ms_halo_reduced'), r.get('ms_halo_chain'), r.get('ms_pair_wall')) for r in d['config']['per_rank']])
    if d.get('developed'): print('   developed ms/step %.4f' % d['developed']['ms_per_step'])
    if d.get('configs2_as_quoted'): print('   configs2_as_quoted ms/step %.4f value %.4g' % (d['configs2_as_quoted']['ms_per_step'], d['configs2_as_quoted']['value']))
except Exception as e:
    print('N=%s weak failed' % N, e)
PY
