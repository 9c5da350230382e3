#!/bin/bash
# Border kernel without its numerically-nothing terms (SFM_K2_SKIP): force / lifecycle / order / step suites (borders held to
# 1e-11 against the reference goldens, enumeration bit-exact, order-independence bitwise), then the default bench line.
O=gpurun_out
timeout 420 python -m pytest tests -m gpu -q > $O/r2_pytest_all_1gpu_v6.log 2>&1; echo "pytest exit $?"; tail -4 $O/r2_pytest_all_1gpu_v6.log
timeout 420 python bench.py > $O/bench_r2_v6_g1.json 2> $O/bench_r2_v6_g1.err; echo "bench exit $?"
SFM_K2_SKIP=0 timeout 200 python bench.py --steps 40 --no-extra --no-cpu-baseline --no-dropin --no-parity > $O/bench_r2_v6_g1_noskip.json 2> $O/bench_r2_v6_g1_noskip.err
timeout 100 python profiles/kernels_alone.py > $O/kernels_alone_r2_v6.log 2>&1; cat $O/kernels_alone_r2_v6.log
for f in bench_r2_v6_g1 bench_r2_v6_g1_noskip; do python - $O/$f.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    r = d['roofline']
    print(sys.argv[1], 'ms/step %.3f' % d['ms_per_step'], 'e2e %.3f' % d['e2e']['ms_per_step'], 'dropin', (d.get('e2e_dropin') or {}).get('ms_per_step'),
          'k1 alone %.3f' % r['ms_per_launch'], 'in step %.3f' % r['ms_per_launch_inside_step'], 'k2 alone %.3f' % d['roofline_k2']['ms_alone'],
          'parity', ((d.get('parity') or {}).get('oracle') or {}).get('worst_err_over_tol'), {k: v.get('ms_per_step') for k, v in d.get('extra', {}).items()}, d['clocks'])
except Exception as e:
    print(sys.argv[1], 'FAILED', e); print(open(sys.argv[1].replace('.json', '.err')).read()[-2000:])
PY
done
