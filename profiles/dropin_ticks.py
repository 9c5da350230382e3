"""Wall time of one tick through the reference-facing Python surface (BASELINE.json configs[0] and [1]).

Drives ``pedestrian_simulation.PedestrianSimulation`` exactly like SimulationRunner.tick does headless:
update_dynamic_obstacles -> tick -> get_new_velocities -> positions advanced by the CARLA stub.  With ``--impl reference``
(build container only: needs /root/reference) the same loop runs on the imported reference for comparison on the same
host.  Prints one JSON line per configuration.
"""
import argparse
import json
import os
import sys
import time
import tomllib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'carla-social-force-model_b200')]


def drive(sim, w, steps, timed_from):
    t0 = None
    for step in range(steps):
        if step == timed_from:
            t0 = time.perf_counter()
        veh = w.vehicles_at(step)
        if veh is not None:
            sim.update_dynamic_obstacles(veh)
        sim.tick(step * w.step_length)
        nv = sim.get_new_velocities()
        sim.peds.state['loc'] += nv['vel'] * w.step_length
    return (time.perf_counter() - t0) / (steps - timed_from)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--configs', default='1,2')
    args = ap.parse_args()
    from sfm_b200 import synth
    with open(os.path.join(ROOT, 'carla-social-force-model_b200', 'config', 'sfm_config.toml'), 'rb') as f:
        cfg = tomllib.load(f)
    for k in (int(x) for x in args.configs.split(',')):
        w = synth.make_config(k)
        steps, warm = (100, 10) if k == 1 else (6, 2)
        if args.impl == 'reference':
            from oracle import ref_loader
            ref = ref_loader.load()
            sim = ref_loader.build_simulation(ref, w, ref_loader.load_config())
            sim.peds.all_states = {}
            sec = drive(sim, w, steps if k == 1 else 3, warm if k == 1 else 1)
        else:
            import pedestrian_simulation
            from ped_mode_manager import PedMode, PedModeManager
            sim = pedestrian_simulation.PedestrianSimulation(list(w.borders), w.section_info(), list(w.static_obstacles),
                                                             cfg, w.step_length, record_states=False)
            modes = [PedModeManager(f'p{i}', float(w.target_speed[i]), PedMode(int(w.mode[i])), 1.0, -1.0) for i in range(w.n)]
            sim.peds.add_pedestrians([f'p{i}' for i in range(w.n)], np.arange(w.n), w.loc, w.vel, w.next_waypoint, modes,
                                     w.radius, w.target_speed)
            sec = drive(sim, w, steps, warm)
        print(json.dumps({'impl': args.impl, 'config': f'cfg{k}', 'n_pedestrians': w.n, 'ms_per_tick': sec * 1e3,
                          'agent_steps_per_s': w.n / sec, 'api': 'PedestrianSimulation.tick (host arrays in and out)'}), flush=True)


if __name__ == '__main__':
    main()
