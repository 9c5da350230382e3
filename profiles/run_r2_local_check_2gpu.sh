#!/bin/bash
# Round-2 check of the staged slot order on TWO GPUs (gpurun --gpus 2): the multi-GPU tests (peer-memory and NCCL exchange,
# bitwise in row order, to rounding with rank-local reordering), then the default bench at 2 GPUs with its parity block --
# which replays the ranks' tick on one GPU under the ranks' slot orders and compares bit for bit.
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py tests/test_gpu_order.py -m gpu -q > $O/r2_pytest_multi_order_2gpu.log 2>&1; echo "pytest exit $?"; tail -4 $O/r2_pytest_multi_order_2gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29751 bench.py --gpus 2 --steps 40 --warmup 3 --no-cpu-baseline > $O/bench_r2_v5_g2.json 2> $O/bench_r2_v5_g2.err; echo "bench exit $?"
python - $O/bench_r2_v5_g2.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    r = d['roofline']
    print('ms/step %.3f' % d['ms_per_step'], 'e2e %.3f' % d['e2e']['ms_per_step'], 'k1 alone %.3f' % r['ms_per_launch'], 'frac %.3f' % r['frac'],
          'local', r['local_tile_pair_fraction']['timed_ticks'], '\nparity', json.dumps(d.get('parity'))[:900],
          '\nextra', {k: (v.get('ms_per_step'), (v.get('parity') or {}).get('ok'), ((v.get('parity') or {}).get('single_gpu_bitwise') or {}).get('identical')) for k, v in d.get('extra', {}).items()}, d['clocks'])
except Exception as e:
    print('FAILED', e); print(open(sys.argv[1].replace('.json', '.err')).read()[-3000:])
PY
