"""Pair-kernel time at N = 65,536 for the four (planar | 3-D) x (radius off | on) code paths."""
import json, os, sys, tomllib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'carla-social-force-model_b200')]
import numpy as np
from sfm_b200 import native, synth
cfg0 = tomllib.load(open(os.path.join(ROOT, 'carla-social-force-model_b200/config/sfm_config.toml'), 'rb'))
n = 65536
for z in (0.0, 0.2):
    for radius in (False, True):
        cfg = dict(cfg0, use_ped_radius=radius)
        w = synth.make_config(5, n=n, z_spread=z)
        ctx = native.Context(0)
        ctx.set_params(native.params_from_config(cfg, 0.05))
        ctx.set_reorder_interval(int(os.environ.get('SFM_REORDER_EVERY', '32')))
        ctx.upload_state(w.loc, w.vel, w.next_waypoint, w.radius, w.target_speed, w.mode)
        out = np.empty((n, 3))
        for _ in range(3): ctx.force(native.PEDESTRIAN, out)
        ctx.reset_stats(); ctx.set_profiling(True)
        for _ in range(8): ctx.force(native.PEDESTRIAN, out)
        s = ctx.stats()
        ms = s['ms_pairs'] / s['pair_launches']
        print(json.dumps({'path': ('3-D' if z else 'planar') + (' + radius' if radius else ''), 'ms': ms,
                          'pair_terms_per_s': s['pair_evaluations'] / s['pair_launches'] / (ms * 1e-3), 'fixup_rows': s['fixup_rows'],
                          'local_tile_pair_fraction': s['local_tile_pairs'] * 65536.0 / max(s['pair_evaluations'], 1)}), flush=True)
        ctx.close()
