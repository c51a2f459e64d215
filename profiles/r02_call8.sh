#!/bin/bash
# Round 2, final scaling points at N GPUs (N = $1): driver-shaped weak run + strong run.
set -u
N=${1:-8}
mkdir -p gpurun_out
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514"
timeout 900 $T bench.py --gpus $N --steps 20 --warmup 5 > $O/r02i_n${N}_weak.json 2> $O/r02i_n${N}_weak.err
timeout 900 $T bench.py --gpus $N --scaling strong --steps 20 --warmup 5 --e2e-steps 2 --no-developed --no-parity > $O/r02i_n${N}_strong.json 2> $O/r02i_n${N}_strong.err
python - $N <<'PY'
import json, sys
N = sys.argv[1]
for tag in ('weak', 'strong'):
    try:
        d = json.load(open('gpurun_out/r02i_n%s_%s.json' % (N, tag)))
        print('N=%s %s ms/step %.4f value %.4g particles %d launches/step %.1f e2e %s' % (N, tag, d['ms_per_step'], d['value'], d['config']['particles'], d['launches_per_step'], d['e2e'].get('ms_per_step')))
        print('   halo', d['config']['halo'])
        p = d['config'].get('multi_gpu_parity')
        if p: print('   parity', p['ok'], max(p['max_scaled_error'].values()))
        print('   per_rank (ms_pair, ms_other, n_real, pairs, sent, reduced, chain, pair_wall)', [(r['ms_pair'], r['ms_other'], int(r['n_real']), int(r['pairs']), r.get('ms_halo_sent'), r.get('ms_halo_reduced'), r.get('ms_halo_chain'), r.get('ms_pair_wall')) for r in d['config']['per_rank']])
        if d.get('developed'): print('   developed ms/step %.4f halo_full %s proactive %s failed %s' % (d['developed']['ms_per_step'], d['developed'].get('halo_full_updates'), d['developed'].get('halo_proactive'), d['developed'].get('halo_deferred_failed')))
        if d.get('configs2_as_quoted'): print('   configs2_as_quoted ms/step %.4f value %.4g' % (d['configs2_as_quoted']['ms_per_step'], d['configs2_as_quoted']['value']))
        for k, v in (d.get('extra') or {}).items():
            print('   extra', k, 'ms/step %.3f value %.4g particles %s' % (v['ms_per_step'], v['value'], v['config'].get('particles')))
    except Exception as e:
        print('N=%s %s failed' % (N, tag), e)
PY
