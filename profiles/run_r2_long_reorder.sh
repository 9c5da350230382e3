#!/bin/bash
# Does the staged order stay good over a long run?  1,000 timed ticks of cfg3 on one GPU with the order rebuilt every 32
# ticks (default) and with the order built once at load and never again; parity block at the END of each run.
O=gpurun_out
timeout 200 python bench.py --steps 1000 --warmup 3 --no-extra --no-cpu-baseline --no-dropin > $O/bench_r2_v5_g1_1000ticks.json 2> $O/bench_r2_v5_g1_1000ticks.err; echo "exit $?"
SFM_REORDER_EVERY=100000000 timeout 200 python bench.py --steps 1000 --warmup 3 --no-extra --no-cpu-baseline --no-dropin > $O/bench_r2_v5_g1_1000ticks_order_once.json 2> $O/bench_r2_v5_g1_1000ticks_order_once.err; echo "exit $?"
for f in bench_r2_v5_g1_1000ticks bench_r2_v5_g1_1000ticks_order_once; do python - $O/$f.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    r = d['roofline']
    print(sys.argv[1], 'ms/step %.3f' % d['ms_per_step'], d['ms_per_step_rank0'], 'k1 alone (after the run) %.3f' % r['ms_per_launch'],
          'local', r['local_tile_pair_fraction'], 'parity', (d.get('parity') or {}).get('oracle', {}).get('worst_err_over_tol'), d['clocks'])
except Exception as e:
    print(sys.argv[1], 'FAILED', e); print(open(sys.argv[1].replace('.json', '.err')).read()[-2000:])
PY
done
