"""Where the drop-in's PedestrianSimulation.tick spends its time at cfg3 (N = 65,536): cProfile over 20 ticks (cumulative
time per function) next to the plain wall time per tick."""
import cProfile, io, os, pstats, sys, time, tomllib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'carla-social-force-model_b200')]
import numpy as np
from sfm_b200 import synth
import pedestrian_simulation
from ped_mode_manager import PedMode, PedModeManager
cfg = tomllib.load(open(os.path.join(ROOT, 'carla-social-force-model_b200/config/sfm_config.toml'), 'rb'))
w = synth.make_config(3)
sim = pedestrian_simulation.PedestrianSimulation(list(w.borders), w.section_info(), list(w.static_obstacles), cfg,
                                                 w.step_length, record_states=False)
names = [f'p_{i}' for i in range(w.n)]
modes = [PedModeManager(names[i], float(w.target_speed[i]), PedMode(int(w.mode[i])), 1.5, 1.0) for i in range(w.n)]
sim.peds.add_pedestrians(names, np.arange(w.n), w.loc, w.vel, w.next_waypoint, modes, w.radius, w.target_speed)


def ticks(k0, count):
    for k in range(k0, k0 + count):
        sim.tick(k * w.step_length)
        nv = sim.get_new_velocities()
        sim.peds.state['loc'] += nv['vel'] * w.step_length


ticks(0, 5)
t0 = time.perf_counter(); ticks(5, 20); print('wall ms per tick (incl. the position stub): %.3f' % ((time.perf_counter() - t0) / 20 * 1e3))
t0 = time.perf_counter()
for k in range(25, 45):
    sim.tick(k * w.step_length)
print('wall ms per tick (tick only): %.3f' % ((time.perf_counter() - t0) / 20 * 1e3))
pr = cProfile.Profile(); pr.enable(); ticks(45, 20); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(28); print(s.getvalue())
