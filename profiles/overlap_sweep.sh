#!/bin/bash
# K1 / K2 co-residency experiments: (auxiliary stream priority, launch order, pair-kernel shared-memory pad) on cfg3.
cat > /tmp/parse_bench.py <<'PY'
import json, sys
lines = [l for l in sys.stdin if l.startswith('{')]
if not lines:
    print('  bench failed:', open('/tmp/err.log').read()[-400:])
else:
    d = json.loads(lines[-1])
    print('  step %.3f ms  k1_in_step %.3f  k2 %.3f  k1_iso %.3f  e2e %.3f' % (
        d['ms_per_step'], d['kernel_ms_per_step']['pairs_k1'], d['kernel_ms_per_step']['segments_cells_k2'],
        d['roofline']['ms_per_launch'], d['e2e']['ms_per_step']))
PY
run() { echo "== $*"; env "$@" python bench.py --no-cpu-baseline --steps 10 2>/tmp/err.log | python /tmp/parse_bench.py; }
run SFM_AUX_PRIORITY=high SFM_K1_FIRST=0 SFM_K1_SMEM_PAD=0
run SFM_AUX_PRIORITY=equal SFM_K1_FIRST=1 SFM_K1_SMEM_PAD=0
run SFM_AUX_PRIORITY=equal SFM_K1_FIRST=1 SFM_K1_SMEM_PAD=40192
run SFM_AUX_PRIORITY=equal SFM_K1_FIRST=1 SFM_K1_SMEM_PAD=74000
run SFM_AUX_PRIORITY=equal SFM_K1_FIRST=1 SFM_K1_SMEM_PAD=20000
run SFM_AUX_PRIORITY=equal SFM_K1_FIRST=0 SFM_K1_SMEM_PAD=40192
run SFM_AUX_PRIORITY=equal SFM_K1_FIRST=0 SFM_K1_SMEM_PAD=0
