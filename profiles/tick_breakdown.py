"""What each force class adds to the cfg3 tick (N = 65,536 + 1.05 M border points + 50 k obstacle points): the tick timed
with subsets of the force switches ([forces] in sfm_config.toml), wall clock over 40 ticks after a warm-up (no L2 flush)."""
import json, os, sys, time, tomllib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'carla-social-force-model_b200')]
import numpy as np
from sfm_b200 import native, synth
cfg = tomllib.load(open(os.path.join(ROOT, 'carla-social-force-model_b200/config/sfm_config.toml'), 'rb'))
w = synth.make_config(3)
names = native.FORCE_CLASSES
subsets = {'all five': names, 'no border': [n for n in names if n != 'border_force'],
           'no static obstacles': [n for n in names if n != 'static_obstacle_force'],
           'pairs + acceleration': ['acceleration_force', 'pedestrian_force'],
           'cell lists + acceleration (no pairs)': [n for n in names if n != 'pedestrian_force']}
for label, on in subsets.items():
    ctx = native.Context(0)
    ctx.set_params(native.params_from_config(cfg, w.step_length, enable={n: (n in on) for n in names}))
    ctx.upload_state(w.loc, w.vel, w.next_waypoint, w.radius, w.target_speed, w.mode)
    ctx.set_borders(w.borders, w.section_center, w.section_length)
    ctx.set_obstacles(native.STATIC_OBSTACLE, [c for c, _ in w.static_obstacles], [r for _, r in w.static_obstacles])
    ctx.step(5); ctx.synchronize()
    t0 = time.perf_counter()
    ctx.step(40); ctx.synchronize()
    print(json.dumps({'forces': label, 'ms_per_tick': (time.perf_counter() - t0) / 40 * 1e3}), flush=True)
    ctx.close()
