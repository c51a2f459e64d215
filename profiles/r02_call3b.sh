#!/bin/bash
# Round 2, GPU call 3b (two B200): N = 2 after moving the boundary CTAs onto the communication
# stream (side by side with the interior CTAs) and folding the dt agreement into the refresh.
set -u
mkdir -p gpurun_out
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
echo "== 2-GPU tests"
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_recut.py tests/test_gpu_rings_multi.py -m gpu -q > $O/r02d_pytest.log 2>&1
tail -3 $O/r02d_pytest.log
for tag in peer nccl; do
  [ $tag = nccl ] && export B200SPH_PEER_SYNC=0 || unset B200SPH_PEER_SYNC
  timeout 400 $T bench.py --gpus 2 --steps 40 --warmup 10 --e2e-steps 3 > $O/r02d_n2_$tag.json 2> $O/r02d_n2_$tag.err
  python - $tag $O/r02d_n2_$tag.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2])); r = d['roofline']
    print('N=2 %s ms/step %.4f value %.4g launches/step %.1f e2e %s' % (sys.argv[1], d['ms_per_step'], d['value'], d['launches_per_step'], d['e2e'].get('ms_per_step')))
    print('   halo', d['config']['halo']); print('   parity', d['config']['multi_gpu_parity']['ok'], d['config']['multi_gpu_parity']['halo'])
    print('   per_rank', d['config']['per_rank']); print('   developed', d.get('developed'))
except Exception as e:
    print('failed', e)
PY
  tail -2 $O/r02d_n2_$tag.err
done
unset B200SPH_PEER_SYNC
echo "== N=2 strong scaling (10 M case)"
timeout 600 $T bench.py --gpus 2 --scaling strong --steps 10 --warmup 5 --e2e-steps 1 --no-developed --no-parity > $O/r02d_n2_strong.json 2> $O/r02d_n2_strong.err
python - $O/r02d_n2_strong.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print('N=2 strong ms/step %.3f value %.4g particles %d' % (d['ms_per_step'], d['value'], d['config']['particles']))
    print('   per_rank', d['config']['per_rank'])
except Exception as e:
    print('failed', e)
PY
tail -2 $O/r02d_n2_strong.err
