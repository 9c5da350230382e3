#!/bin/bash
# Grid depth of the pair kernel (SFM_K1_TARGET_CTAS; nsplit = ceil(target / own tiles), capped at half the tile count): the
# shorter a CTA (fewer partner tiles each), the smaller the tail and the imbalance between CTAs of 3 and 4 tiles -- against
# the per-CTA prologue / epilogue.  cfg3 tick on one GPU.
O=gpurun_out
for ctas in 9472 14208 16384 18944 23680 28416 37888; do
  SFM_K1_TARGET_CTAS=$ctas python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-extra --no-parity --no-dropin 2>/dev/null | python -c "
import sys, json
l = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('target CTAs %-6s tick %.3f ms  e2e %.3f  k1 alone %.3f  k1 in step %.3f' % ('$ctas', l['ms_per_step'], l['e2e']['ms_per_step'], l['roofline']['ms_per_launch'], l['roofline']['ms_per_launch_inside_step']))" >> $O/k1_target_ctas_r2_v5.log
done
cat $O/k1_target_ctas_r2_v5.log
