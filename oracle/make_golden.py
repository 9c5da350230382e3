"""Generate tests/golden/*.npz by running the imported, unmodified reference (build container only).

TEST INFRASTRUCTURE.  Usage:  python -m oracle.make_golden [--only cfg1|cfg2|cfg2traj|lifecycle|csv]
The inputs are regenerated from ``sfm_b200.synth`` by seed; each file stores a SHA-256 of the input bytes so that a
drift of the generator is detected instead of silently comparing against stale outputs.
"""
from __future__ import annotations

import argparse
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'carla-social-force-model_b200')]

from oracle import ref_loader          # noqa: E402
from sfm_b200 import synth             # noqa: E402

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def workload_digest(w):
    h = hashlib.sha256()
    for a in (w.loc, w.vel, w.next_waypoint, w.radius, w.target_speed, w.mode):
        h.update(np.ascontiguousarray(a).tobytes())
    for b in w.borders:
        h.update(np.ascontiguousarray(b).tobytes())
    if w.section_center is not None:
        h.update(np.ascontiguousarray(w.section_center).tobytes())
        h.update(np.ascontiguousarray(w.section_length).tobytes())
    for c, r in w.static_obstacles:
        h.update(np.ascontiguousarray(c).tobytes())
        h.update(np.ascontiguousarray(r).tobytes())
    veh = w.vehicles_at(3)
    if veh is not None:
        for ring in veh[5]:
            h.update(np.ascontiguousarray(ring).tobytes())
    return h.hexdigest()


def golden_cfg1(ref, cfg):
    w = synth.make_config(1)
    out = ref_loader.run_ticks(ref, w, cfg, 100, record_forces=True)
    keep = [0, 1, 10, 50, 99]
    np.savez_compressed(os.path.join(GOLDEN, 'cfg1_trajectory.npz'), digest=workload_digest(w), loc=out['loc'],
                        vel=out['vel'], force_steps=np.array(keep),
                        **{f'F_{k}': v[keep] for k, v in out['forces'].items()})


def golden_cfg2(ref, cfg):
    for use_radius in (False, True):
        for z_spread in (0.0, 0.2):
            w = synth.make_config(2, z_spread=z_spread)
            c = dict(cfg, use_ped_radius=use_radius)
            sim = ref_loader.build_simulation(ref, w, c)
            sim.update_dynamic_obstacles(w.vehicles_at(0))
            t0 = time.time()
            forces = {name: f.get_force(sim.peds) for name, f in sim.forces.items()}
            print(f'cfg2 radius={use_radius} z={z_spread}: {time.time() - t0:.1f}s', flush=True)
            tag = f"r{int(use_radius)}_z{int(z_spread > 0)}"
            np.savez_compressed(os.path.join(GOLDEN, f'cfg2_forces_{tag}.npz'), digest=workload_digest(w),
                                **{f'F_{k}': v for k, v in forces.items()})


CFG2_TRAJECTORY_STEPS = (1, 5, 10)


def golden_cfg2_trajectory(ref, cfg):
    """Ten ticks of the reference's own loop at N = 4,096 (dense all-pairs regime, all five forces, vehicles moving):
    positions and velocities of every second pedestrian after ticks 1, 5 and 10 (about 3 minutes, 4.4 GB)."""
    w = synth.make_config(2)
    t0 = time.time()
    out = ref_loader.run_ticks(ref, w, cfg, max(CFG2_TRAJECTORY_STEPS), record_forces=False)
    print(f'cfg2 trajectory: {time.time() - t0:.1f}s', flush=True)
    keep = list(CFG2_TRAJECTORY_STEPS)
    np.savez_compressed(os.path.join(GOLDEN, 'cfg2_trajectory.npz'), digest=workload_digest(w), steps=np.array(keep),
                        rows=np.arange(0, w.n, 2), loc=out['loc'][keep][:, ::2], vel=out['vel'][keep][:, ::2])


def lifecycle_digest(w, life):
    h = hashlib.sha256(workload_digest(w).encode())
    for route in life.routes:
        for wp, crossing in route:
            h.update(np.ascontiguousarray(wp).tobytes())
            h.update(bytes([int(crossing)]))
    for a in (life.crossing_speed_factor, life.crossing_safety_margin, life.idle, w.veh_center, w.veh_vel, w.veh_yaw):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


LIFECYCLE_STEPS = 140


def golden_lifecycle(ref, cfg):
    """Mode machines + gap acceptance + waypoint hand-over, run by the reference's own classes (ref_loader.run_lifecycle)."""
    w, life = synth.make_lifecycle()
    out = ref_loader.run_lifecycle(ref, w, life, cfg, LIFECYCLE_STEPS)
    np.savez_compressed(os.path.join(GOLDEN, 'lifecycle.npz'), digest=lifecycle_digest(w, life), **out)


def pack_ragged(hist, n):
    """Ragged per-tick histories of a run with despawns -> dense arrays: alive mask, modes (255 = gone), final state."""
    ticks = len(hist['ids'])
    alive = np.zeros((ticks, n), dtype=bool)
    mode = np.full((ticks, n), 255, dtype=np.uint8)
    for k in range(ticks):
        alive[k, hist['ids'][k]] = True
        mode[k, hist['ids'][k]] = hist['mode'][k]
    return dict(alive=alive, mode=mode, ids_final=np.asarray(hist['ids'][-1]), loc_final=np.asarray(hist['loc'][-1]),
                vel_final=np.asarray(hist['vel'][-1]), wp_final=np.asarray(hist['wp'][-1]))


def golden_lifecycle_despawn(ref, cfg):
    """The same scenario with despawn_on_arrival (run_simulation.py:38,127-132): the crowd shrinks from 48 to 10."""
    w, life = synth.make_lifecycle()
    out = ref_loader.run_lifecycle(ref, w, life, cfg, LIFECYCLE_STEPS, despawn=True)
    np.savez_compressed(os.path.join(GOLDEN, 'lifecycle_despawn.npz'), digest=lifecycle_digest(w, life),
                        **pack_ragged(out, w.n))


def golden_lifecycle_spawn(ref, cfg):
    """Spawning (two late waves, pedestrian_simulation.py:99-100) and despawning in one run: the crowd goes 34 -> 33 -> 18."""
    import dataclasses
    w, life = synth.make_lifecycle(spawn_late=14)
    life = dataclasses.replace(life, despawn_on_arrival=True)
    out = ref_loader.run_lifecycle(ref, w, life, cfg, 100, despawn=True)
    np.savez_compressed(os.path.join(GOLDEN, 'lifecycle_spawn.npz'), digest=lifecycle_digest(w, life),
                        spawn_tick=life.spawn_tick, **pack_ragged(out, w.n))


def golden_output_csv(ref_dir):
    """The four CSV files the reference's own OutputGenerator writes for a tiny recorded scene (output_generator.py)."""
    import importlib.util
    import tempfile
    import types
    spec = importlib.util.spec_from_file_location('_ref_output_generator', os.path.join(ref_dir, 'output_generator.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    scene = synth.make_output_scene()
    ped_sim = types.SimpleNamespace(peds=types.SimpleNamespace(all_states=scene['ped_states']),
                                    all_dyn_obs_states=scene['veh_states'], static_obstacles=scene['static_obstacles'],
                                    borders=scene['borders'])
    files = {}
    with tempfile.TemporaryDirectory() as tmp:
        gen = mod.OutputGenerator(ped_sim, tmp, 'golden')
        gen.generate_ped_csv(); gen.generate_veh_csv(); gen.generate_borders_csv(); gen.generate_obstacles_csv()
        for name in ('pedestrian.csv', 'vehicle.csv', 'borders.csv', 'obstacles.csv'):
            with open(os.path.join(gen.output_dir, name), 'rb') as f:
                files[name.replace('.', '_')] = np.frombuffer(f.read(), dtype=np.uint8)
    np.savez_compressed(os.path.join(GOLDEN, 'output_csv.npz'), **files)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--only', default=None)
    args = ap.parse_args()
    ref, cfg = ref_loader.load(), ref_loader.load_config()
    os.makedirs(GOLDEN, exist_ok=True)
    if args.only in (None, 'cfg1'):
        golden_cfg1(ref, cfg)
    if args.only in (None, 'cfg2'):
        golden_cfg2(ref, cfg)
    if args.only in (None, 'cfg2traj'):
        golden_cfg2_trajectory(ref, cfg)
    if args.only in (None, 'lifecycle'):
        golden_lifecycle(ref, cfg)
        golden_lifecycle_despawn(ref, cfg)
        golden_lifecycle_spawn(ref, cfg)
    if args.only in (None, 'csv'):
        golden_output_csv(ref_loader.REFERENCE_DIR)


if __name__ == '__main__':
    main()
