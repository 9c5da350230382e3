#!/bin/bash
# K1s variant sweep: rows per thread (separate builds under build/) x grid depth, at N = 65,536 (flat crowd).
for ir in 1 2 4; do
  for ctas in 2368 9472 37888; do
    SFM_LIB=$PWD/build/libsfm_ir$ir.so SFM_K1_TARGET_CTAS=$ctas python - <<PY
import os, sys, tomllib
sys.path[:0] = ['.', 'carla-social-force-model_b200']
import numpy as np
from sfm_b200 import native, synth
cfg = tomllib.load(open('carla-social-force-model_b200/config/sfm_config.toml', 'rb'))
n = 65536
w = synth.make_config(5, n=n)
ctx = native.Context(0)
ctx.set_params(native.params_from_config(cfg, 0.05))
ctx.upload_state(w.loc, w.vel, w.next_waypoint, w.radius, w.target_speed, w.mode)
out = np.empty((n, 3))
for _ in range(2): ctx.force(native.PEDESTRIAN, out)
ctx.reset_stats(); ctx.set_profiling(True)
for _ in range(5): ctx.force(native.PEDESTRIAN, out)
s = ctx.stats()
print(f"IR=$ir target_ctas=$ctas: {s['ms_pairs'] / s['pair_launches']:.3f} ms")
PY
  done
done
