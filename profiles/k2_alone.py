"""Cell-list kernels alone at cfg3 (one CTA per pedestrian group): event-timed border / static-obstacle force launches.
Used under ncu (`-k regex:k2_segments`) for the source-level counters of K2."""
import json, os, sys, tomllib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'carla-social-force-model_b200')]
import numpy as np
from sfm_b200 import native, synth
cfg = tomllib.load(open(os.path.join(ROOT, 'carla-social-force-model_b200/config/sfm_config.toml'), 'rb'))
w = synth.make_config(3)
ctx = native.Context(0)
ctx.set_params(native.params_from_config(cfg, w.step_length))
ctx.upload_state(w.loc, w.vel, w.next_waypoint, w.radius, w.target_speed, w.mode)
ctx.set_borders(w.borders, w.section_center, w.section_length)
ctx.set_obstacles(native.STATIC_OBSTACLE, [c for c, _ in w.static_obstacles], [r for _, r in w.static_obstacles])
out = np.empty((w.n, 3))
reps = int(os.environ.get('K2_REPS', '5'))
res = {}
for cls, name in ((native.BORDER, 'border'), (native.STATIC_OBSTACLE, 'static')):
    ctx.force(cls, out)
    ctx.reset_stats(); ctx.set_profiling(True)
    for _ in range(reps):
        ctx.force(cls, out)
    s = ctx.stats()
    ctx.set_profiling(False)
    res[name] = {'ms_segments': s['ms_segments'] / max(reps, 1), 'ms_cells': s['ms_cells'] / max(reps, 1),
                 'checksum': float(np.abs(out).sum())}
print(json.dumps(res))
