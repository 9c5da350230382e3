#!/bin/bash
# Tick time (bench.py, cfg3) for every build/libsfm_*.so variant: the pair kernel shares the SMs with the cell-list kernels
# inside a tick, so the variant is chosen on ms_per_step, not on the kernel alone.   usage: profiles/bench_variants.sh [steps]
cd "$(dirname "$0")/.."
steps=${1:-30}
for lib in build/libsfm_*.so; do
  name=$(basename $lib .so); name=${name#libsfm_}
  SFM_LIB=$PWD/$lib python bench.py --steps $steps --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
l = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-16s tick %.3f ms  e2e %.3f ms  k1 alone %.3f  k1 in step %.3f  k2 span %.3f  frac %.3f' % ('$name', l['ms_per_step'], l['e2e']['ms_per_step'], l['roofline']['ms_per_launch'], l['roofline']['ms_per_launch_inside_step'], l['kernel_ms_per_step']['segments_cells_k2'], l['roofline']['frac']))"
done
