#!/bin/bash
# Round 2, GPU call 2 (TWO B200): the 2-GPU tests, the peer protocol against the NCCL path at
# N = 2, the k_pair_list variants at N = 1, compute-sanitizer.
#   gpurun --gpus 2 --timeout 1500 -- 'bash profiles/r02_call2.sh'
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $O/r02b_gpu.txt 2>&1
nvidia-smi topo -m >> $O/r02b_gpu.txt 2>&1

echo "== pytest -m gpu (2 GPUs visible)" | tee $O/r02b_pytest.log
timeout 900 python -m pytest tests -m gpu -q -rs --durations=10 >> $O/r02b_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $O/r02b_pytest.log
tail -12 $O/r02b_pytest.log

B="python bench.py --steps 40 --warmup 10 --no-cpu --e2e-steps 3 --no-extras --no-developed"
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    r = d['roofline']
    print('%s ms/step %.4f pair %.4f nnps %.4f other %.4f launches/step %.1f e2e %s value %.3e' % (
        sys.argv[1], d['ms_per_step'], r['avg_launch_ms'], r['ms_nnps_per_step'], r['ms_other_per_step'],
        d.get('launches_per_step', -1), d['e2e'].get('ms_per_step'), d['value']))
    h = d['config'].get('halo')
    if h: print('   halo', h)
    p = d['config'].get('multi_gpu_parity')
    if p: print('   parity', p['ok'], p['max_scaled_error'], p['halo'])
    if d['config'].get('per_rank'): print('   per_rank', d['config']['per_rank'])
except Exception as e:
    print(sys.argv[1], 'failed', e)
PY
}
for v in "7 1" "8 1" "7 0" "6 0"; do
  set -- $v
  B200SPH_PAIR_MINB=$1 B200SPH_PAIR_SPEC=$2 timeout 300 $B > $O/r02b_n1_minb$1_spec$2.json 2> $O/r02b_n1_minb$1_spec$2.err
  show "N=1 MINB=$1 SPEC=$2" $O/r02b_n1_minb$1_spec$2.json
done

T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
echo "== N=2 peer protocol"
timeout 400 $T bench.py --gpus 2 --steps 40 --warmup 10 --e2e-steps 3 > $O/r02b_n2_peer.json 2> $O/r02b_n2_peer.err
show "N=2 peer" $O/r02b_n2_peer.json
tail -3 $O/r02b_n2_peer.err
echo "== N=2 NCCL scalars, no overlap (B200SPH_PEER_SYNC=0)"
B200SPH_PEER_SYNC=0 timeout 400 $T bench.py --gpus 2 --steps 40 --warmup 10 --e2e-steps 3 --no-developed > $O/r02b_n2_nccl.json 2> $O/r02b_n2_nccl.err
show "N=2 nccl" $O/r02b_n2_nccl.json
echo "== N=2 MINB=8"
B200SPH_PAIR_MINB=8 timeout 400 $T bench.py --gpus 2 --steps 40 --warmup 10 --e2e-steps 3 --no-developed --no-parity > $O/r02b_n2_peer8.json 2> $O/r02b_n2_peer8.err
show "N=2 peer MINB=8" $O/r02b_n2_peer8.json

echo "== compute-sanitizer memcheck: smoke()"
timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02b_memcheck_smoke.log 2>&1
tail -4 $O/r02b_memcheck_smoke.log
echo "== compute-sanitizer racecheck: smoke()"
timeout 400 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02b_racecheck_smoke.log 2>&1
tail -4 $O/r02b_racecheck_smoke.log
echo "== compute-sanitizer memcheck: 2-GPU slab test (peer protocol)"
timeout 500 compute-sanitizer --tool memcheck --target-processes all --print-limit 20 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > $O/r02b_memcheck_multi.log 2>&1
tail -6 $O/r02b_memcheck_multi.log
ls -la $O | grep r02b
