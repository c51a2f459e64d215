#!/bin/bash
# Round 2, GPU call 1 (one B200): the whole GPU test suite with the xfail marks gone, the new
# full-size parity test, the rewritten k_pair_list at 6 / 7 / 8 CTAs per SM, the fused stage
# kernel on / off, and the ncu captures VERDICT r01 asked for.
#   gpurun --timeout 1500 -- 'bash profiles/r02_call1.sh'
# Everything lands in gpurun_out/r02a_*; summaries are copied to profiles/ by hand.
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r02a_gpu.txt 2>&1

echo "== pytest -m gpu" | tee $O/r02a_pytest.log
timeout 900 python -m pytest tests -m gpu -q -x --durations=15 >> $O/r02a_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $O/r02a_pytest.log
tail -5 $O/r02a_pytest.log

B="python bench.py --steps 40 --warmup 10 --no-cpu --e2e-steps 3"
for m in 6 7 8; do
  echo "== bench MINB=$m"
  B200SPH_PAIR_MINB=$m timeout 300 $B > $O/r02a_bench_minb$m.json 2> $O/r02a_bench_minb$m.err
  python - <<EOF
import json
try:
    d = json.load(open('$O/r02a_bench_minb$m.json'))
    r = d['roofline']
    print('MINB=$m ms/step %.4f pair %.4f ms nnps %.4f other %.4f launches %s' % (d['ms_per_step'], r['avg_launch_ms'], r['ms_nnps_per_step'], r['ms_other_per_step'], d.get('gpu_launches')))
except Exception as e:
    print('MINB=$m failed', e)
EOF
done
echo "== bench FUSE=0"
B200SPH_FUSE=0 timeout 300 $B > $O/r02a_bench_nofuse.json 2> $O/r02a_bench_nofuse.err
python - <<EOF
import json
try:
    d = json.load(open('$O/r02a_bench_nofuse.json'))
    r = d['roofline']
    print('FUSE=0 ms/step %.4f pair %.4f ms nnps %.4f other %.4f launches %s' % (d['ms_per_step'], r['avg_launch_ms'], r['ms_nnps_per_step'], r['ms_other_per_step'], d.get('gpu_launches')))
except Exception as e:
    print('FUSE=0 failed', e)
EOF
echo "== bench default (with cpu baseline), 100 steps"
timeout 400 python bench.py > $O/r02a_bench_default.json 2> $O/r02a_bench_default.err
tail -c 600 $O/r02a_bench_default.json

# ---- ncu: launch list of the dam-break step, then full captures --------------------------
NCU="ncu --clock-control none"
echo "== ncu launch list (dam break)"
timeout 400 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/r02a_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --e2e-steps 1 > $O/r02a_ncu_launch.log 2>&1
echo "== ncu full: k_pair_list, k_stage_pack"
timeout 500 $NCU --set full --import-source on -k regex:'k_pair_list|k_stage_pack' -s 8 -c 3 -o $O/r02a_pair \
    python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > $O/r02a_ncu_pair.log 2>&1
echo "== ncu full: k_list_build<false>"
timeout 400 $NCU --set full --import-source on -k regex:k_list_build -c 1 -o $O/r02a_listbuild \
    python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > $O/r02a_ncu_lb.log 2>&1
echo "== ncu full: rings (k_solid_pass1/2)"
timeout 500 $NCU --set full --import-source on -k regex:'k_solid_pass' -s 8 -c 2 -o $O/r02a_solid \
    python bench.py --workload rings --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > $O/r02a_ncu_solid.log 2>&1
echo "== ncu full: taylor-green (k_tvf_pass1/2, k_list_build<true>)"
timeout 600 $NCU --set full --import-source on -k regex:'k_tvf_pass|k_list_build' -s 6 -c 3 -o $O/r02a_tvf \
    python bench.py --workload taylor_green --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > $O/r02a_ncu_tvf.log 2>&1
echo "== rings + taylor-green bench lines"
timeout 300 python bench.py --workload rings --steps 20 --warmup 5 --no-cpu --e2e-steps 2 > $O/r02a_bench_rings.json 2> $O/r02a_bench_rings.err
timeout 300 python bench.py --workload taylor_green --steps 20 --warmup 5 --no-cpu --e2e-steps 2 > $O/r02a_bench_tg.json 2> $O/r02a_bench_tg.err
ls -la $O | tail -30
