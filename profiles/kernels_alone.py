"""The three hot kernels alone at cfg3 (N = 65,536 + 1.05 M border points + 50 k obstacle points), one launch each after a
warm-up: the target of `ncu --set full -k "regex:k1_sym_pairs|k2_segments"` (profiles/r2_ncu_*.csv) and of the
launch-list pass.  KA_WARM=0 skips the warm-up launches (so that -c 3 captures exactly one launch per kernel)."""
import json, os, sys, tomllib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'carla-social-force-model_b200')]
import numpy as np
from sfm_b200 import native, synth
cfg = tomllib.load(open(os.path.join(ROOT, 'carla-social-force-model_b200/config/sfm_config.toml'), 'rb'))
w = synth.make_config(3)
ctx = native.Context(0)
ctx.set_params(native.params_from_config(cfg, w.step_length))
ctx.set_reorder_interval(int(os.environ.get('SFM_REORDER_EVERY', '32')))     # staged along the Hilbert curve, like Engine
ctx.upload_state(w.loc, w.vel, w.next_waypoint, w.radius, w.target_speed, w.mode)
ctx.set_borders(w.borders, w.section_center, w.section_length)
ctx.set_obstacles(native.STATIC_OBSTACLE, [c for c, _ in w.static_obstacles], [r for _, r in w.static_obstacles])
out = np.empty((w.n, 3))
res = {}
for cls, name in ((native.BORDER, 'border'), (native.STATIC_OBSTACLE, 'static'), (native.PEDESTRIAN, 'pairs')):
    for _ in range(int(os.environ.get('KA_WARM', '1'))):
        ctx.force(cls, out)
    ctx.reset_stats(); ctx.set_profiling(True)
    ctx.force(cls, out)
    s = ctx.stats()
    ctx.set_profiling(False)
    res[name] = {'ms': s['ms_pairs'] + s['ms_segments'] + s['ms_cells'], 'checksum': float(np.abs(out).sum())}
print(json.dumps(res))
