#!/bin/bash
# Tick time (bench.py, cfg3) for every build/libsfm_*.so variant x SFM_K2_PERSIST setting: the pair kernel shares the SMs
# with the cell-list kernels inside a tick, so variants are chosen on ms_per_step, not on a kernel alone.
# usage: profiles/bench_variants.sh [steps] ["persist values"]
cd "$(dirname "$0")/.."
steps=${1:-30}
persists=${2:-2}
libs=$(ls build/libsfm_*.so 2>/dev/null)
[ -z "$libs" ] && libs=carla-social-force-model_b200/sfm_b200/libsfm_b200.so      # no variants built: the in-tree library
for lib in $libs; do
  name=$(basename $lib .so); name=${name#libsfm_}
  for p in $persists; do
  SFM_K2_PERSIST=$p SFM_LIB=$PWD/$lib python bench.py --steps $steps --warmup 3 --no-cpu-baseline --no-extra --no-parity --no-dropin 2>/dev/null | python -c "
import sys, json
l = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-14s persist=%s tick %.3f ms  e2e %.3f ms  k1 alone %.3f  k1 in step %.3f  k2 alone %.3f  k2 span %.3f' % ('$name', '$p', l['ms_per_step'], l['e2e']['ms_per_step'], l['roofline']['ms_per_launch'], l['roofline']['ms_per_launch_inside_step'], l['roofline_k2']['ms_alone'], l['kernel_ms_per_step']['segments_cells_k2']))"
  done
done
