"""Device time of the lifecycle kernels on the cfg4 crowd: N = 262,144 pedestrians, 2,048 vehicles.

Each kernel is timed by the library's own CUDA events (sfm_stats.ms_lifecycle) over repeated launches and set against
its algorithmic bytes and the measured HBM bandwidth (MEASURED_PEAKS.json).  Prints one JSON object.
"""
import json
import os
import sys
import tomllib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'carla-social-force-model_b200')]
from sfm_b200 import native, synth          # noqa: E402


def timed(ctx, fn, reps):
    fn()
    ctx.synchronize()
    ctx.reset_stats()
    ctx.set_profiling(True)
    for _ in range(reps):
        fn()
    s = ctx.stats()
    ctx.set_profiling(False)
    return s['ms_lifecycle'] / reps


def main():
    n = int(os.environ.get('LIFE_N', '262144'))
    with open(os.path.join(ROOT, 'carla-social-force-model_b200', 'config', 'sfm_config.toml'), 'rb') as f:
        cfg = tomllib.load(f)
    peak = 6551.0
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        peak = float(json.load(open(p)).get('hbm_gbs', peak))
    w = synth.make_config(4, n=n)
    rng = np.random.default_rng(7)
    ctx = native.Context(0)
    ctx.set_params(native.params_from_config(cfg, w.step_length))
    mode = rng.choice([1, 2, 4], size=n, p=[0.7, 0.1, 0.2]).astype(np.uint8)        # 20 % waiting at the kerb
    wp = w.loc + np.column_stack((rng.normal(0.0, 7.0, (n, 2)), np.zeros(n)))           # a road crossing: the far kerb ~10 m away
    ctx.upload_state(w.loc, w.vel, wp, w.radius, w.target_speed, mode)
    ctx.set_mode_machines(w.target_speed, 1.5 * w.target_speed, rng.uniform(0.5, 2.5, n))
    ctx.set_vehicles(w.veh_center, w.veh_yaw, w.veh_vel, w.veh_extent, w.veh_resolution)
    v = len(w.veh_center)
    _, rings = ctx.download_vehicles()
    n_points = sum(len(r) for r in rings)
    routes_off = np.arange(n + 1, dtype=np.int64) * 2
    wps = np.repeat(w.next_waypoint, 2, axis=0) + rng.normal(0, 3.0, (2 * n, 3))
    ctx.set_routes_csr(routes_off, wps, rng.integers(0, 2, 2 * n).astype(np.uint8), 2.0, fused=False)
    out = {'n_pedestrians': n, 'n_vehicles': v, 'ring_points': n_points, 'hbm_peak_gbs': peak, 'kernels': {}}

    def entry(name, ms, bytes_, note):
        gbs = bytes_ / (ms * 1e-3) / 1e9
        out['kernels'][name] = {'ms': ms, 'algorithmic_bytes': bytes_, 'gb_per_s': gbs, 'frac_of_measured_hbm': gbs / peak,
                                'note': note}

    def tick():
        ctx.update_targets(mode=mode)               # put the waiting pedestrians back (not timed: ms_lifecycle only)
        ctx.tick_modes(0.0)
    checking = int((mode == 4).sum())
    ms = timed(ctx, tick, 10)
    entry('k4_tick_modes', ms, n * (32 + 32 + 8 + 1 + 1) + checking * (16 + 24) + v * 32,
          f'{checking} pedestrians x {v} vehicles gap acceptance = {checking * v / (ms * 1e-3) / 1e9:.2f} G segment tests/s; '
          'compute/latency bound, the bytes are the per-row state it must touch')
    ms = timed(ctx, ctx.advance_waypoints, 20)
    entry('k4_advance_waypoints', ms, n * (32 + 16 + 4 + 4), 'arrival test; hand-over rows add 24+24+1+8 B each')
    ms = timed(ctx, lambda: ctx.advance_vehicles(w.step_length), 20)
    entry('k5_advance+rings (+rebin, timed separately as cells)', ms, v * (16 * 2 + 16 + 16 + 8) + n_points * 16,
          f'{n_points} ring points regenerated; replaces a {n_points * 16 / 1e6:.1f} MB host upload per tick; launch-latency bound')
    ctx.record_begin(32)
    ms = timed(ctx, lambda: (ctx.record_begin(32), ctx.record_frame(0.0)), 20)
    entry('k6_record_frame', ms, n * (64 + 1 + 32 + 1), '(x, y, vx, vy, mode) snapshot')
    print(json.dumps(out))


if __name__ == '__main__':
    main()
