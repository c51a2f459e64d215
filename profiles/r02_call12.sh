#!/bin/bash
# Round 2: the GPU suite with the Group-control and generic-equation (NVRTC) tests, then the
# driver-shaped N = 1 bench line without the CPU leg.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests -m gpu -q --durations=8 > $O/r02m_pytest.log 2>&1
echo "pytest rc=$?"; tail -25 $O/r02m_pytest.log
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --e2e-steps 3 --no-extras > $O/r02m_bench.json 2> $O/r02m_bench.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r02m_bench.json'))
    r = d['roofline']
    print('N=1 ms/step %.4f value %.4g pair %.4f ms frac %.3f other %.4f launches/step %s developed %.4f' % (
        d['ms_per_step'], d['value'], r['avg_launch_ms'], r['frac'], r['ms_other_per_step'], d['launches_per_step'], d['developed']['ms_per_step']))
except Exception as e:
    print('bench failed', e)
PY
