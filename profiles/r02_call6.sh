#!/bin/bash
# Round 2, GPU call 6 (one B200): early skin adaptation; driver-shaped and BASELINE-protocol runs;
# launch list of the final step.
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest -m gpu" | tee $O/r02g_pytest.log
timeout 900 python -m pytest tests -m gpu -q -x >> $O/r02g_pytest.log 2>&1
tail -3 $O/r02g_pytest.log
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2])); r = d['roofline']
    print('%s ms/step %.4f pair %.4f value %.4g e2e %.3f launches/step %.1f nnps %s' % (sys.argv[1], d['ms_per_step'], r['avg_launch_ms'], d['value'], d['e2e']['ms_per_step'], d['launches_per_step'], r['nnps']))
    if d.get('developed'): print('   developed ms/step %.4f builds %d proactive %d failed %d ms/rebuild %s' % (d['developed']['ms_per_step'], d['developed']['full_builds'], d['developed']['proactive_builds'], d['developed']['deferred_failed'], d['developed']['ms_per_rebuild']))
except Exception as e:
    print(sys.argv[1], 'failed', e)
PY
}
timeout 400 python bench.py --steps 20 --warmup 5 --no-extras > $O/r02g_20_5.json 2> $O/r02g_20_5.err; show "20/5" $O/r02g_20_5.json
timeout 400 python bench.py --steps 100 --warmup 50 --no-extras --no-cpu > $O/r02g_100_50.json 2> $O/r02g_100_50.err; show "100/50" $O/r02g_100_50.json
B200SPH_SKIN_ADAPT=0 timeout 400 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu --no-developed > $O/r02g_noadapt.json 2> $O/r02g_noadapt.err; show "20/5 skin fixed 0.1" $O/r02g_noadapt.json
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02g_ref.json 2> $O/r02g_ref.err; python -c "
import json; d=json.load(open('$O/r02g_ref.json')); print('reference arm', d['value'], d['ms_per_step'], d['config']['workload'], d['warmup'], d['steps'], d['cpu_baseline']['cores'])"
echo "== ncu launch list"
timeout 400 ncu --clock-control none --metrics gpu__time_duration.sum -c 300 --csv --log-file $O/r02g_launches.csv \
    python bench.py --steps 4 --warmup 8 --no-cpu --e2e-steps 1 --no-extras --no-developed > $O/r02g_ncu_launch.log 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(l for l in open('gpurun_out/r02g_launches.csv') if l.startswith('"')))
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Value')
d=collections.OrderedDict()
for r in rows[1:]:
    d.setdefault(r[ki].split('(')[0],[]).append(float(r[mi].replace(',',''))/1000.0)
for n,v in sorted(d.items(), key=lambda kv:-sum(kv[1]))[:12]:
    print('%-40s %4d avg %8.2f us'%(n[:40],len(v),sum(v)/len(v)))
PY
