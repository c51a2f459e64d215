#!/bin/bash
# Round 2: the GPU suite with the EDAC solid-wall tests; N = 1 dam-break line and the Taylor-Green
# line (the EDAC kernels without walls must be where they were: 4.34 ms / step).
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests -m gpu -q > $O/r02n_pytest.log 2>&1
echo "pytest rc=$?"; tail -8 $O/r02n_pytest.log
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --e2e-steps 3 --no-extras --no-developed > $O/r02n_bench.json 2> $O/r02n_bench.err
timeout 200 python bench.py --workload taylor_green --steps 20 --warmup 5 --no-cpu --e2e-steps 2 > $O/r02n_tg.json 2> $O/r02n_tg.err
python - <<'PY'
import json
for f in ('r02n_bench', 'r02n_tg'):
    try:
        d = json.load(open('gpurun_out/%s.json' % f))
        r = d['roofline']
        print('%s ms/step %.4f value %.4g kernel %s %.4f ms frac %.3f' % (f, d['ms_per_step'], d['value'], r['kernel'], r['avg_launch_ms'], r['frac']))
    except Exception as e:
        print(f, 'failed', e)
PY
