"""[historical: the SFM_K1_IR / SFM_K1_MINB knobs of the v2 row kernel were removed after this sweep picked IR=2, MINB=5;
results in tune_k1_r1.log]  K1 tuning sweep (rows per thread, min CTAs/SM, grid depth) at N = 65,536 -- prints ms per launch for each variant."""
import itertools
import os
import sys
import tomllib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'carla-social-force-model_b200')]
import numpy as np                      # noqa: E402
from sfm_b200 import native, synth      # noqa: E402

cfg = tomllib.load(open(os.path.join(ROOT, 'carla-social-force-model_b200', 'config', 'sfm_config.toml'), 'rb'))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
w = synth.make_config(5, n=n)
variants = list(itertools.product((2, 1), (5, 6, 8, 10), (148 * 4 * 16, 148 * 4 * 4, 148 * 4 * 64)))
for ir, minb, ctas in variants:
    if (ir == 2 and minb == 10) or (ir == 1 and minb in (5, 6)):
        continue
    os.environ.update(SFM_K1_IR=str(ir), SFM_K1_MINB=str(minb), SFM_K1_TARGET_CTAS=str(ctas))
    ctx = native.Context(0)
    ctx.set_params(native.params_from_config(cfg, 0.05))
    ctx.upload_state(w.loc, w.vel, w.next_waypoint, w.radius, w.target_speed, w.mode)
    out = np.empty((n, 3))
    for _ in range(2):
        ctx.force(native.PEDESTRIAN, out)
    ctx.reset_stats()
    ctx.set_profiling(True)
    for _ in range(5):
        ctx.force(native.PEDESTRIAN, out)
    s = ctx.stats()
    ms = s['ms_pairs'] / s['pair_launches']
    print(f'IR={ir} MINB={minb:2d} target_ctas={ctas:6d}: {ms:8.3f} ms  {n * (n - 1) / ms / 1e9:8.1f} Gpairs/s  '
          f'frac58={n * (n - 1) / ms * 1e3 * 58 / 37.22496e12:.3f}', flush=True)
    ctx.close()
