#!/bin/bash
# Final state of round 2 (v5 + deeper pair-kernel grid): whole -m gpu suite and the default bench line on one GPU.
O=gpurun_out
timeout 300 python -m pytest tests -m gpu -q > $O/r2_pytest_all_1gpu_final.log 2>&1; echo "pytest exit $?"; tail -3 $O/r2_pytest_all_1gpu_final.log
timeout 300 python bench.py > $O/bench_r2_final_g1.json 2> $O/bench_r2_final_g1.err; echo "bench exit $?"
python - $O/bench_r2_final_g1.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    r = d['roofline']
    print('ms/step %.3f' % d['ms_per_step'], 'e2e %.3f' % d['e2e']['ms_per_step'], 'dropin', (d.get('e2e_dropin') or {}).get('ms_per_step'),
          'k1 alone %.3f' % r['ms_per_launch'], 'frac %.3f' % r['frac'], 'in step %.3f' % r['ms_per_launch_inside_step'],
          'parity', ((d.get('parity') or {}).get('oracle') or {}).get('worst_err_over_tol'), {k: v.get('ms_per_step') for k, v in d.get('extra', {}).items()}, d['clocks'])
except Exception as e:
    print('FAILED', e); print(open(sys.argv[1].replace('.json', '.err')).read()[-2000:])
PY
