#!/bin/bash
# Round 2, v5 kernel: compile-time variants of the pair kernel (build/libsfm_*.so, built by profiles/build_k1s_variants.sh)
# and two run-time settings (rebuild interval of the staged order, persistent cell-list CTAs) on the cfg3 tick.
O=gpurun_out
bash profiles/bench_variants.sh 30 "2" > $O/bench_variants_r2_v5.log 2>&1
for every in 8 128; do
  SFM_REORDER_EVERY=$every python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-extra --no-parity --no-dropin 2>/dev/null | python -c "
import sys, json
l = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('reorder every %-4s tick %.3f ms  k1 alone %.3f  local share %.3f' % ('$every', l['ms_per_step'], l['roofline']['ms_per_launch'], l['roofline']['local_tile_pair_fraction']['timed_ticks']))" >> $O/bench_variants_r2_v5.log
done
for p in 1 3; do
  SFM_K2_PERSIST=$p python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-extra --no-parity --no-dropin 2>/dev/null | python -c "
import sys, json
l = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('in-tree library persist=%s tick %.3f ms  k1 in step %.3f' % ('$p', l['ms_per_step'], l['roofline']['ms_per_launch_inside_step']))" >> $O/bench_variants_r2_v5.log
done
for ctas in 4736 18944; do
  SFM_K1_TARGET_CTAS=$ctas python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-extra --no-parity --no-dropin 2>/dev/null | python -c "
import sys, json
l = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('in-tree library target CTAs %-6s tick %.3f ms  k1 alone %.3f  k1 in step %.3f' % ('$ctas', l['ms_per_step'], l['roofline']['ms_per_launch'], l['roofline']['ms_per_launch_inside_step']))" >> $O/bench_variants_r2_v5.log
done
cat $O/bench_variants_r2_v5.log
