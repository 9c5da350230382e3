#!/bin/bash
# Persistent cell-list kernels beside the pair kernel: CTAs per SM x stream priority, cfg3 tick.
cat > /tmp/parse_bench.py <<'PY'
import json, sys
lines = [l for l in sys.stdin if l.startswith('{')]
if not lines:
    print('  bench failed:', open('/tmp/err.log').read()[-400:])
else:
    d = json.loads(lines[-1])
    print('  step %.3f ms  k1_in_step %.3f  k2_span %.3f  k1_iso %.3f  e2e %.3f' % (
        d['ms_per_step'], d['kernel_ms_per_step']['pairs_k1'], d['kernel_ms_per_step']['segments_cells_k2'],
        d['roofline']['ms_per_launch'], d['e2e']['ms_per_step']))
PY
run() { echo "== $*"; env "$@" python bench.py --no-cpu-baseline --steps 10 2>/tmp/err.log | python /tmp/parse_bench.py; }
run SFM_K2_PERSIST=0
run SFM_K2_PERSIST=1
run SFM_K2_PERSIST=2
run SFM_K2_PERSIST=3
run SFM_K2_PERSIST=4
run SFM_K2_PERSIST=2 SFM_AUX_PRIORITY=equal
run SFM_K2_PERSIST=3 SFM_AUX_PRIORITY=equal
run SFM_K2_PERSIST=2 SFM_AUX_PRIORITY=low
