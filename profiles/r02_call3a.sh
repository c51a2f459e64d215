#!/bin/bash
# Round 2, GPU call 3a (one B200): BASELINE configs[2]'s 10 M-particle case (dx = 0.00407,
# 11.3 M particles) on ONE GPU, cell rows row-major vs along the Z-curve of (cy, cz):
# ms / launch, L1 / L2 hit rates of k_pair_list, + the new GPU tests.
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest (new tests)" | tee $O/r02c_pytest.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "zorder or fused or monaghan" >> $O/r02c_pytest.log 2>&1
tail -3 $O/r02c_pytest.log
B="python bench.py --dx 0.00407 --steps 10 --warmup 6 --no-cpu --e2e-steps 1 --no-developed --no-extras"
for z in 0 1; do
  echo "== 11.3 M particles on one GPU, B200SPH_ZORDER=$z"
  B200SPH_ZORDER=$z timeout 600 $B > $O/r02c_11m_z$z.json 2> $O/r02c_11m_z$z.err
  python - $O/r02c_11m_z$z.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); r = d['roofline']
    print('ms/step %.3f pair %.4f ms nnps %.3f other %.3f particles %d pairs/step %.4g value %.4g' % (
        d['ms_per_step'], r['avg_launch_ms'], r['ms_nnps_per_step'], r['ms_other_per_step'], d['config']['particles'], d['config']['pairs_per_step'], d['value']))
except Exception as e:
    print('failed', e)
PY
  tail -2 $O/r02c_11m_z$z.err
done
M="gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__throughput.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active"
for z in 0 1; do
  echo "== ncu k_pair_list at 11.3 M, B200SPH_ZORDER=$z"
  B200SPH_ZORDER=$z timeout 900 ncu --clock-control none --metrics $M -k regex:'k_pair_list|k_stage_pack|k_list_build' -s 4 -c 5 --csv --log-file $O/r02c_ncu_11m_z$z.csv \
     python bench.py --dx 0.00407 --steps 1 --warmup 3 --no-cpu --e2e-steps 1 --no-developed --no-extras > $O/r02c_ncu_11m_z$z.log 2>&1
  grep -c k_pair_list $O/r02c_ncu_11m_z$z.csv
done
# the same comparison at configs[1] (1.22 M: records L2-resident)
for z in 0 1; do
  echo "== 1.22 M, B200SPH_ZORDER=$z"
  B200SPH_ZORDER=$z timeout 300 python bench.py --steps 40 --warmup 10 --no-cpu --e2e-steps 2 --no-developed --no-extras > $O/r02c_1m_z$z.json 2> $O/r02c_1m_z$z.err
  python - $O/r02c_1m_z$z.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); r = d['roofline']
    print('ms/step %.4f pair %.4f ms' % (d['ms_per_step'], r['avg_launch_ms']))
except Exception as e:
    print('failed', e)
PY
done
# developed regime + extras at N = 1 (the driver's default invocation shape, shorter)
echo "== default-shaped run (developed + extras)"
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02c_default.json 2> $O/r02c_default.err
python - $O/r02c_default.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print('ms/step %.4f value %.4g e2e %s' % (d['ms_per_step'], d['value'], d['e2e']))
    print('developed', d.get('developed'))
    for k, v in (d.get('extra') or {}).items():
        print(k, 'ms/step %.3f value %.4g roofline %.3f' % (v['ms_per_step'], v['value'], v['roofline']['frac']))
except Exception as e:
    print('failed', e)
PY
tail -3 $O/r02c_default.err
