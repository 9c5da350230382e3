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


class _OpaqueMode(PedModeManager):
    """A subclass is not a *stock* PedModeManager: a table containing one takes the generic per-object path."""
    __slots__ = ()


def _lifecycle_sim(w, life, cfg, record_states, opaque=False):
    reset_session()
    sim = pedestrian_simulation.PedestrianSimulation(list(w.borders), w.section_info(), list(w.static_obstacles), cfg,
                                                     w.step_length, record_states=record_states)
    cls = _OpaqueMode if opaque else PedModeManager
    for i in range(w.n):
        mode = cls(f'p_{i}', float(w.target_speed[i]), PedMode(int(w.mode[i])), float(life.crossing_speed_factor[i]),
                   float(life.crossing_safety_margin[i]))
        if life.idle[i]:
            mode.set_mode(PedMode.IDLE)
        sim.spawn_pedestrian((f'p_{i}', i, w.loc[i], w.vel[i], w.next_waypoint[i], mode, w.radius[i], w.target_speed[i]))
    return sim


def _drive(sim, w, life, steps):
    """SimulationRunner.tick with CARLA stubbed (run_simulation.py:77-132): vehicles -> tick -> move -> waypoint hand-over."""
    routes = {f'p_{i}': list(r) for i, r in enumerate(life.routes)}
    log = []
    for k in range(steps):
        sim.update_dynamic_obstacles(w.vehicles_at(k))
        sim.tick(k * w.step_length)
        nv = sim.get_new_velocities()
        assert np.shares_memory(nv, sim.peds.state)
        for name in sim.get_arrived_peds(life.waypoint_threshold):          # arrival at the positions the forces saw
            if routes[name]:
                sim.peds.update_next_waypoint(name, routes[name].pop(0))
        sim.peds.state['loc'] += nv['vel'] * w.step_length
        log.append((sim.peds.mode_codes().copy(), sim.peds.state['vel'].copy(), sim.peds.state['target_speed'].copy(),
                    np.array([float(m.next_mode_time) for m in sim.peds.state['mode']])))
    return log


def test_resident_columnar_and_generic_ticks_agree(sfm_config):
    """The three ways through PedestrianSimulation.tick (device-resident K4a, columnar host machines, per-object generic)
    produce the same modes, target speeds, wake-up times and velocities tick by tick on the lifecycle scene: idle
    pedestrians waking up, gap acceptance against moving vehicles, hand-overs requesting crossings."""
    w, life = synth.make_lifecycle()
    steps = 80
    generic = _drive(_lifecycle_sim(w, life, sfm_config, record_states=False, opaque=True), w, life, steps)
    columnar_sim = _lifecycle_sim(w, life, sfm_config, record_states=True)
    columnar = _drive(columnar_sim, w, life, steps)
    resident_sim = _lifecycle_sim(w, life, sfm_config, record_states=False)
    resident = _drive(resident_sim, w, life, steps)
    assert resident_sim.peds.mode_table() is not None and len(columnar_sim.get_states()) == steps
    seen = set()
    for k in range(steps):
        for other, name in ((columnar, 'columnar'), (resident, 'resident')):
            for a, b, what in zip(generic[k], other[k], ('modes', 'velocities', 'target speeds', 'wake-up times')):
                assert np.array_equal(a, b), f'{name} path: {what} differ at tick {k}'
        seen.update(generic[k][0].tolist())
    assert seen == {0, 1, 2, 3, 4}                               # every mode occurred
    snap = columnar_sim.get_states()[0.0]
    assert snap['mode'].dtype == object and len(snap) == w.n           # recorded like pedestrian_state.py:100-104


def test_resident_tick_keeps_sets_and_rows_on_device(sfm_config):
    """Steady-state resident ticks upload the pedestrian table and nothing else: no parameter, point-set or full-state
    upload, no interpreter loop (the launch count per tick is constant and small)."""
    w = synth.make_config(2)
    sim = build_sim(w, sfm_config, record_states=False)
    from sfm_b200.session import get_session
    sim.tick(0.0)
    ctx = get_session().ctx
    sim.tick(0.05)
    ctx.reset_stats()
    for k in range(5):
        sim.tick(0.1 + 0.05 * k)
    s = ctx.stats()
    assert s['steps'] == 5 and s['launches'] <= 5 * 24, s
    # the velocities are the reference's: same state through the per-class force objects and the host-side update
    state = sim.peds.state.copy()
    F = sum(f.get_force(sim.peds) for f in sim.forces.values())
    sim.peds.state['vel'] = state['vel']
    sim.calculate_new_velocities(F)
    want = sim.get_new_velocities()['vel'].copy()
    sim.peds.state['vel'] = state['vel']
    sim.tick(1.0)
    assert np.abs(sim.peds.state['vel'] - want).max() < 1e-6


def test_calculate_new_velocities_on_device_equals_numpy(sfm_config):
    """calculate_new_velocities(force) for a caller-composed force array runs on the device (sfm_apply_force) in numpy's
    operation order: bit-identical to pedestrian_simulation.py:117-124 / stateutils.py:18-23 evaluated by the oracle."""
    from oracle import sfm_oracle as O
    w = synth.make_config(1)
    sim = build_sim(w, sfm_config, record_states=False)
    rng = np.random.default_rng(3)
    force = rng.normal(0.0, 30.0, size=(w.n, 3))
    force[5] = 0.0
    sim.peds.state['vel'][5] = 0.0                                   # zero speed: the divide-by-one branch
    want = O.new_velocities(sim.peds.state['vel'].copy(), force, sim.peds.state['target_speed'].copy(), w.step_length)
    sim.calculate_new_velocities(force)
    got = sim.get_new_velocities()
    assert np.shares_memory(got, sim.peds.state)
    np.testing.assert_array_equal(got['vel'], want)


def test_device_identity_check_notices_replaced_objects(sfm_config, monkeypatch):
    """After its first tick the resident path no longer walks the `mode` column on the host: the device compares the
    object pointers (sfm_tick_records, identity 'check').  An object replaced IN PLACE -- same array, same length -- must
    still be noticed: the tick then hands the table back untouched and the full path rebuilds the machines, so the run
    equals one whose table held that object from the start; steady ticks never call the host-side comparison."""
    from sfm_b200 import native
    # rows staged in row order: the two runs below are compared bit for bit, and the replaced object triggers one extra
    # full upload -- which would rebuild the staged order (csrc/k8_order.cuh) at a different tick than the plain run does
    monkeypatch.setenv('SFM_REORDER_EVERY', '0')
    w = synth.make_config(2, n=4096)                       # >= 4096 rows: the 2-D DMA write-back path
    calls = []
    real = native.column_equal
    monkeypatch.setattr(native, 'column_equal', lambda *a, **k: calls.append(1) or real(*a, **k))

    def run(replace_at):
        sim = build_sim(w, sfm_config, record_states=False)
        out = []
        for k in range(6):
            if k == replace_at:
                old = sim.peds.state['mode'][7]
                new = PedModeManager(old.ped_name, 0.5 * old.initial_target_speed, PedMode.WALKING_SIDEWALK, 1.0, -1.0)
                sim.peds.state['mode'][7] = new                # in place: the array object and its length stay the same
            sim.tick(0.05 * k)
            out.append((sim.peds.state['vel'].copy(), sim.peds.state['target_speed'].copy()))
            sim.peds.state['loc'] += sim.get_new_velocities()['vel'] * w.step_length
        return out

    calls.clear()
    plain = run(None)
    assert len(calls) <= 1, calls                           # steady ticks: no host pass over the column
    swapped = run(3)
    for k in range(3):
        assert np.array_equal(plain[k][0], swapped[k][0])
    assert swapped[3][1][7] == 0.5 * plain[3][1][7]          # the new object's speed governs from its first tick on
    assert not np.array_equal(plain[3][0][7], swapped[3][0][7])
    others = np.arange(w.n) != 7
    assert np.array_equal(plain[3][0][others], swapped[3][0][others])      # same positions, only row 7's clamp differs
    assert np.array_equal(swapped[5][1][others], plain[5][1][others])
