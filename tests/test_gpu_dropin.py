"""The drop-in Python surface (forces / pedestrian_state / pedestrian_simulation) driven like the reference drives it."""
import numpy as np
import pytest

import forces
import pedestrian_simulation
from ped_mode_manager import PedMode, PedModeManager
from sfm_b200 import synth
from sfm_b200.session import reset_session
from tests import golden_util as G
from tests.gpu_util import assert_forces_close

pytestmark = pytest.mark.gpu


def build_sim(w, cfg, record_states=True):
    """Spawn the workload the way PedSpawnManager does: one 8-tuple per pedestrian (pedestrian_spawner.py:240-241)."""
    reset_session()
    sim = pedestrian_simulation.PedestrianSimulation(list(w.borders), w.section_info(), list(w.static_obstacles), cfg,
                                                     w.step_length, record_states=record_states)
    for i in range(w.n):
        name = f'p{i}'
        mode = PedModeManager(name, float(w.target_speed[i]), PedMode(int(w.mode[i])), 1.0, -1.0)
        sim.spawn_pedestrian((name, i, w.loc[i], w.vel[i], w.next_waypoint[i], mode, w.radius[i], w.target_speed[i]))
    return sim


def test_cfg1_tick_loop_against_reference_golden(sfm_config):
    """The reference's own tick loop shape: update_dynamic_obstacles -> tick -> get_new_velocities -> CARLA stub."""
    w = synth.make_config(1)
    g = G.load('cfg1_trajectory.npz', w)
    sim = build_sim(w, sfm_config)
    assert list(sim.forces) == ['acceleration_force', 'pedestrian_force', 'border_force', 'static_obstacle_force',
                                'dynamic_obstacle_force']
    keep = list(g['force_steps'])
    for step in range(100):
        sim.update_dynamic_obstacles(w.vehicles_at(step))
        if step in keep[:3]:
            for name, f in sim.forces.items():
                got, want = f.get_force(sim.peds), g[f'F_{name}'][keep.index(step)]
                assert got.shape == (w.n, 3) and got.dtype == np.float64
                if step == 0 and name != 'pedestrian_force':
                    np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-11, err_msg=name)
                elif step == 0:
                    assert_forces_close(got, want, name=name, atol=2e-5)
        sim.tick(step * w.step_length)
        nv = sim.get_new_velocities()
        assert np.shares_memory(nv, sim.peds.state)               # the view aliasing callers rely on
        sim.peds.state['loc'] += nv['vel'] * w.step_length
    dev = np.linalg.norm(sim.peds.state['loc'] - g['loc'][100], axis=1)
    assert dev.max() <= 1e-2 and np.median(dev) <= 1e-4, (dev.max(), np.median(dev))
    assert len(sim.get_states()) == 100 and len(sim.all_dyn_obs_states) == 100


def test_force_switches_and_mode_mask(sfm_config):
    cfg = dict(sfm_config, forces=dict(acceleration_force=True, border_force=True))
    w = synth.make_config(1)
    sim = build_sim(w, cfg, record_states=False)
    assert list(sim.forces) == ['acceleration_force', 'border_force']
    border = sim.forces['border_force'].get_force(sim.peds)
    crossing = w.mode == 2
    assert crossing.any() and not border[crossing].any() and border[~crossing].any()
    sim.tick(0.0)
    assert sim.get_states() == {}
    with pytest.raises(KeyError):
        forces.PedestrianForce(0.05, {})                             # mandatory section, like forces.py:66


def test_spawn_despawn_between_ticks(sfm_config):
    w = synth.make_config(1)
    sim = build_sim(w, sfm_config, record_states=False)
    sim.update_dynamic_obstacles(w.vehicles_at(0))
    sim.tick(0.0)
    sim.destroy_pedestrian('p3')
    sim.destroy_pedestrian('p60')
    assert sim.peds.size() == w.n - 2
    sim.tick(0.05)
    assert sim.get_new_velocities()['vel'].shape == (w.n - 2, 3)
    arrived = sim.get_arrived_peds(1e9)
    assert len(arrived) == w.n - 2
    empty = pedestrian_simulation.PedestrianSimulation([], np.empty((0, 2), dtype=object), [], sfm_config, 0.05)
    empty.tick(0.0)                                                  # no pedestrians: early-out
    assert empty.get_new_velocities() is None
