#!/bin/bash
# Round 2, GPU call 4 (one B200): the new list builder (tests, developed regime, ncu), arithmetic
# cell offsets in every list consumer.
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest -m gpu" | tee $O/r02e_pytest.log
timeout 900 python -m pytest tests -m gpu -q -x >> $O/r02e_pytest.log 2>&1
tail -3 $O/r02e_pytest.log
echo "== default-shaped run"
timeout 600 python bench.py --steps 40 --warmup 10 > $O/r02e_default.json 2> $O/r02e_default.err
python - $O/r02e_default.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); r = d['roofline']
    print('ms/step %.4f pair %.4f value %.4g e2e %.3f launches/step %.1f' % (d['ms_per_step'], r['avg_launch_ms'], d['value'], d['e2e']['ms_per_step'], d['launches_per_step']))
    print('developed', d.get('developed'))
    for k, v in (d.get('extra') or {}).items():
        print(k, 'ms/step %.3f value %.4g roofline %.3f nnps %.3f other %.3f' % (v['ms_per_step'], v['value'], v['roofline']['frac'], v['roofline']['ms_nnps_per_step'], v['roofline']['ms_other_per_step']))
except Exception as e:
    print('failed', e)
PY
tail -3 $O/r02e_default.err
echo "== ncu: k_list_build<false> (dam break), k_list_build<true> (taylor-green), k_pair_list"
timeout 500 ncu --clock-control none --set full --import-source on -k regex:'k_list_build|k_pair_list' -c 3 -o $O/r02e_lb \
    python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 --no-developed --no-extras > $O/r02e_ncu_lb.log 2>&1
timeout 500 ncu --clock-control none --set full -k regex:'k_list_build' -c 1 -o $O/r02e_lbp \
    python bench.py --workload taylor_green --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > $O/r02e_ncu_lbp.log 2>&1
ls -la $O | grep r02e
