#!/bin/bash
# Round-2 check of the run-local pair path + staged slot order on one GPU: the whole -m gpu suite, then the default
# bench configuration (short) with the reordering on and off.  Every command is bounded by its own timeout.
O=gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > $O/r2_pytest_local_1gpu.log 2>&1; echo "pytest exit $?"; tail -5 $O/r2_pytest_local_1gpu.log
timeout 200 python bench.py --steps 40 --warmup 3 --no-extra --no-cpu-baseline --no-dropin > $O/bench_r2_v5_quick.json 2> $O/bench_r2_v5_quick.err; echo "bench exit $?"
SFM_REORDER_EVERY=0 timeout 200 python bench.py --steps 40 --warmup 3 --no-extra --no-cpu-baseline --no-dropin --no-parity > $O/bench_r2_v5_quick_noreorder.json 2> $O/bench_r2_v5_quick_noreorder.err; echo "bench (row order) exit $?"
for f in bench_r2_v5_quick bench_r2_v5_quick_noreorder; do python - $O/$f.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    r = d['roofline']
    print(sys.argv[1], 'ms/step %.3f' % d['ms_per_step'], 'e2e %.3f' % d['e2e']['ms_per_step'], 'k1 alone %.3f' % r['ms_per_launch'],
          'frac %.3f' % r['frac'], 'local', r.get('local_tile_pair_fraction'), 'parity', (d.get('parity') or {}).get('oracle'), d['clocks'])
except Exception as e:
    print(sys.argv[1], 'FAILED', e); print(open(sys.argv[1].replace('.json', '.err')).read()[-2000:])
PY
done
