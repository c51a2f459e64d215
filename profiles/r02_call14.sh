#!/bin/bash
# Round 2, last GPU call: the whole GPU suite (Group control, NVRTC generic equations, EDAC walls /
# external flow, laminar viscosity) and the driver-shaped N = 1 line.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests -m gpu -q > $O/r02o_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 $O/r02o_pytest.log
timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu --e2e-steps 3 --no-extras > $O/r02o_bench.json 2> $O/r02o_bench.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r02o_bench.json'))
    r = d['roofline']
    print('N=1 ms/step %.4f value %.4g pair %.4f ms frac %.3f host_loop %.4f remeasured %s developed %.4f e2e %.3f' % (
        d['ms_per_step'], d['value'], r['avg_launch_ms'], r['frac'], d['host_loop_ms_per_step'], d['remeasured'],
        d['developed']['ms_per_step'], d['e2e']['ms_per_step']))
except Exception as e:
    print('bench failed', e)
PY
