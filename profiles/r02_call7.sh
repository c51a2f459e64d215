#!/bin/bash
# Round 2, GPU call 7 (one B200): EDAC / elastic passes on the shared list walk; MINB 7 vs 8 again.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $O/r02h_pytest.log 2>&1; tail -2 $O/r02h_pytest.log
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2])); r = d['roofline']
    print('%s ms/step %.4f pair %.4f value %.4g frac %.3f nnps %.3f other %.3f e2e %.3f' % (sys.argv[1], d['ms_per_step'], r['avg_launch_ms'], d['value'], r['frac'], r['ms_nnps_per_step'], r['ms_other_per_step'], d['e2e']['ms_per_step']))
except Exception as e:
    print(sys.argv[1], 'failed', e)
PY
}
timeout 300 python bench.py --workload taylor_green --steps 40 --warmup 10 --no-cpu --e2e-steps 3 > $O/r02h_tg.json 2> $O/r02h_tg.err; show "taylor_green" $O/r02h_tg.json
timeout 300 python bench.py --workload rings --steps 40 --warmup 10 --no-cpu --e2e-steps 3 > $O/r02h_rings.json 2> $O/r02h_rings.err; show "rings" $O/r02h_rings.json
for m in 7 8; do
B200SPH_PAIR_MINB=$m timeout 300 python bench.py --steps 40 --warmup 10 --no-cpu --e2e-steps 2 --no-extras --no-developed > $O/r02h_minb$m.json 2> $O/r02h_minb$m.err; show "dam break MINB=$m" $O/r02h_minb$m.json
done
echo "== ncu: tvf / solid passes after the change"
timeout 500 ncu --clock-control none --set full -k regex:'k_solid_pass' -s 8 -c 2 -o $O/r02h_solid python bench.py --workload rings --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > $O/r02h_ncu_solid.log 2>&1
timeout 500 ncu --clock-control none --set full -k regex:'k_tvf_pass' -s 6 -c 2 -o $O/r02h_tvf python bench.py --workload taylor_green --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > $O/r02h_ncu_tvf.log 2>&1
for r in r02h_solid r02h_tvf; do ncu -i $O/$r.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for r in rows[2:]:
    print(r[h.index('Kernel Name')][:30], r[h.index('gpu__time_duration.sum')], 'regs', r[h.index('launch__registers_per_thread')], 'issue', r[h.index('smsp__issue_active.avg.pct_of_peak_sustained_active')], 'l1tex', r[h.index('l1tex__throughput.avg.pct_of_peak_sustained_active')], 'warps', r[h.index('sm__warps_active.avg.pct_of_peak_sustained_active')], 'inst', r[h.index('smsp__inst_executed.sum')])
"; done
