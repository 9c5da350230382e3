"""-m gpu: the device lifecycle through the C ABI against the oracle and the reference goldens:
mode machines + gap acceptance (K4a), arrival test + waypoint hand-over fused into K3 (K4b), vehicle rings generated on
the device (K5), the frame recorder (K6)."""
import os

import numpy as np
import pytest

from oracle import lifecycle_oracle as LO
from oracle import sfm_oracle as O
from oracle.make_golden import LIFECYCLE_STEPS
from sfm_b200 import native, synth
from sfm_b200.headless import HeadlessRunner
from tests.golden_util import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def scenario():
    return synth.make_lifecycle()


@pytest.fixture(scope='module')
def golden():
    return np.load(os.path.join(GOLDEN, 'lifecycle.npz'))


def test_free_running_device_loop_matches_reference_golden(scenario, golden, sfm_config):
    """All five forces + lifecycle on the device for 140 ticks; only the pair force is float32, so positions drift by
    <= 1e-3 m while every discrete decision (mode, hand-over tick, target speed) equals the reference's."""
    w, life = scenario
    run = HeadlessRunner(sfm_config, w, life)
    for k in range(LIFECYCLE_STEPS):
        run.tick()
        s = run.snapshot()
        assert np.array_equal(s['mode'], golden['mode'][k + 1]), f'modes differ after tick {k}'
        assert np.array_equal(s['wp'], golden['wp'][k + 1]), f'waypoints differ after tick {k}'
        assert np.array_equal(s['target_speed'], golden['target_speed'][k]), f'applied target speeds differ at tick {k}'
        assert np.array_equal(s['mode_target_speed'], golden['mode_speed'][k + 1])
    err = np.abs(s['loc'] - golden['loc'][-1])
    assert err.max() < 0.1 and np.median(err) < 1e-4       # 140 chaotic steps: a few close encounters amplify 1e-6
    remaining = np.array([len(r) for r in life.routes]) - s['cursor']
    assert np.array_equal(remaining, golden['remaining'][-1])
    counters = run.ctx.lifecycle_counters()
    m = golden['mode'].astype(int)
    assert counters['handovers'] == int((np.diff(golden['remaining'], axis=0) != 0).sum())
    assert counters['idle_wakeups'] == int(((m[:-1] == 0) & (m[1:] != 0)).sum())
    assert counters['finished'] == int(s['finished'].sum()) > 0


def test_teacher_forced_decisions_equal_reference(scenario, golden, sfm_config):
    """Same loop, but the kinematics of every tick are the reference's own (sfm_update_kinematics): inputs identical =>
    every decision identical, including the borderline ones."""
    w, life = scenario
    run = HeadlessRunner(sfm_config, w, life)
    for k in range(LIFECYCLE_STEPS):
        run.ctx.update_kinematics(golden['loc'][k], golden['vel'][k])
        run.tick()
        s = run.snapshot()
        assert np.array_equal(s['mode'], golden['mode'][k + 1]) and np.array_equal(s['wp'], golden['wp'][k + 1])
        assert np.abs(s['vel'] - golden['vel'][k + 1]).max() < 2e-5        # one float32 pair-force step


def test_gap_acceptance_many_vehicles(sfm_config):
    """Every pedestrian CHECKING_TRAFFIC against 300 vehicles (more than one shared-memory tile): device decisions ==
    oracle decisions, pedestrian by pedestrian."""
    rng = np.random.default_rng(17)
    w = synth.make_config(2, n=1024)
    v = 300
    centres, vels = rng.uniform(0, w.side, (v, 2)), rng.normal(0, 6, (v, 2))
    vels[rng.random(v) < 0.1] = 0.0
    extents = np.tile([2.4, 1.0], (v, 1))
    margin = rng.uniform(-0.3, 2.0, w.n)
    wp = w.loc.copy()
    wp[:, :2] += rng.normal(0, 4.0, (w.n, 2))
    ctx = native.Context(0)
    ctx.set_params(native.params_from_config(sfm_config, w.step_length))
    mode = np.full(w.n, 4, dtype=np.uint8)
    ctx.upload_state(w.loc, w.vel, wp, w.radius, w.target_speed, mode)
    ctx.set_mode_machines(w.target_speed, 1.5 * w.target_speed, margin)
    ctx.set_traffic(centres, vels, extents)
    ctx.tick_modes(0.0)
    got = ctx.download_modes()
    want = np.array([LO.check_traffic(w.loc[i], wp[i], 1.5 * w.target_speed[i], margin[i], centres, vels, extents)
                     for i in range(w.n)])
    assert np.array_equal(got['mode'] == 2, want)
    assert 0.05 < want.mean() < 0.95
    assert np.array_equal(got['mode_target_speed'], np.where(want, 1.5 * w.target_speed, w.target_speed))
    assert np.array_equal(got['target_speed'], w.target_speed)            # applied before the machines changed
    # no vehicles: everybody crosses (pedestrian_simulation.py:68-73)
    ctx.update_targets(mode=mode)
    ctx.set_traffic([], [], [])
    ctx.tick_modes(0.05)
    assert (ctx.download_mode_codes() == 2).all()


def test_device_vehicle_rings_and_forces(sfm_config):
    """sfm_set_vehicles / sfm_advance_vehicles: ring points equal the restated obstacles.py:269-281 to 1e-12, the
    neighbour enumeration on them is bit-exact, and the force equals the host-ring path on the very same points."""
    w = synth.make_config(4, n=8192)
    v = len(w.veh_center)
    ctx = native.Context(0)
    ctx.set_params(native.params_from_config(sfm_config, w.step_length))
    ctx.upload_state(w.loc, w.vel, w.next_waypoint, w.radius, w.target_speed, w.mode)
    ctx.set_vehicles(w.veh_center, w.veh_yaw, w.veh_vel, w.veh_extent, w.veh_resolution)
    centre = w.veh_center.copy()
    for k in range(3):
        if k:
            ctx.advance_vehicles(w.step_length)
            centre = centre + w.veh_vel * w.step_length
        got_c, rings = ctx.download_vehicles()
        assert np.array_equal(got_c, centre)
        for j in range(v):
            want = LO.ellipse_ring(centre[j], w.veh_yaw[j], w.veh_extent[j, 0], w.veh_extent[j, 1], w.veh_resolution)
            assert rings[j].shape == want.shape and np.abs(rings[j] - want).max() < 1e-12
        f_dev = ctx.force(native.DYNAMIC_OBSTACLE)
        pairs_dev = ctx.enumerate_pairs(native.DYNAMIC_OBSTACLE)
        p = O.moussaid_params(sfm_config['dynamic_obstacle_force'], O.OBSTACLE_DEFAULTS)
        f_ref, pairs_ref = O.obstacle_force(w.loc, w.vel, w.radius, got_c, rings, w.veh_vel, p, False, return_pairs=True)
        assert np.array_equal(pairs_dev, pairs_ref)
        assert np.abs(f_dev - f_ref).max() < 1e-10
        host = native.Context(0)
        host.set_params(native.params_from_config(sfm_config, w.step_length))
        host.upload_state(w.loc, w.vel, w.next_waypoint, w.radius, w.target_speed, w.mode)
        host.set_obstacles(native.DYNAMIC_OBSTACLE, got_c, rings, w.veh_vel)
        f_host = host.force(native.DYNAMIC_OBSTACLE)
        # same points, same kernel; after an advance the device set keeps its initial grid, the host path re-plans it, so
        # the (cell, index) summation order may differ in the last bits
        assert np.array_equal(f_host, f_dev) if k == 0 else np.abs(f_host - f_dev).max() < 1e-13
        host.close()
    assert len(pairs_ref) > 1000


def test_standalone_waypoint_advance_and_recorder(scenario, sfm_config, tmp_path):
    """sfm_advance_waypoints as its own call == the oracle's hand-over; the recorder's frames == the device state at the
    recorded ticks, and the CSV written from them has the reference's schema."""
    w, life = scenario
    ctx = native.Context(0)
    ctx.set_params(native.params_from_config(sfm_config, w.step_length))
    ctx.upload_state(w.loc, w.vel, w.next_waypoint, w.radius, w.target_speed, w.mode)
    ctx.set_mode_machines(w.target_speed, life.crossing_speed_factor * w.target_speed, life.crossing_safety_margin)
    ctx.set_routes(life.routes, 2.5, fused=False)
    machines = LO.Machines.create(w.target_speed, w.mode, life.crossing_speed_factor, life.crossing_safety_margin)
    wp, cursor, finished = w.next_waypoint.copy(), np.zeros(w.n, dtype=np.int64), np.zeros(w.n, dtype=bool)
    ctx.record_begin(4)
    for k in range(4):
        ctx.record_frame(0.05 * k)
        ctx.advance_waypoints()
        LO.advance_waypoints(machines, w.loc, wp, life.routes, cursor, finished, 2.5)
        got_cursor, got_finished, got_wp = ctx.download_routes()
        assert np.array_equal(got_cursor, cursor) and np.array_equal(got_finished, finished)
        assert np.array_equal(got_wp, wp) and np.array_equal(ctx.download_mode_codes(), machines.mode)
    assert cursor.sum() > 10 and finished.any()
    times, xyv, mode = ctx.download_frames()
    assert np.array_equal(times, 0.05 * np.arange(4)) and xyv.shape == (4, w.n, 4)
    assert np.array_equal(xyv[0], np.column_stack((w.loc[:, :2], w.vel[:, :2]))) and np.array_equal(mode[0], w.mode)
    with pytest.raises(native.SfmError):
        ctx.record_frame(1.0)                                             # buffer full: fails loudly
    import types
    from output_generator import OutputGenerator
    sim = types.SimpleNamespace(peds=types.SimpleNamespace(all_states={}), all_dyn_obs_states={}, static_obstacles=[],
                                borders=[])
    gen = OutputGenerator(sim, str(tmp_path), 'dev')
    gen.generate_ped_csv(device_frames=(times, xyv, mode), ped_ids=np.arange(w.n))
    lines = open(os.path.join(gen.output_dir, 'pedestrian.csv')).read().splitlines()
    assert lines[0] == 'ped_id,frame,time,x,y,v_x,v_y,mode' and len(lines) == 1 + 4 * w.n
    assert lines[1].split(',')[:3] == ['0', '0', '0.0']


def test_headless_loop_with_device_vehicles(scenario, sfm_config):
    """The fully device-resident loop: vehicles advanced and their rings regenerated on the device every tick, gap
    acceptance against that set.  Oracle: the same loop with centres accumulated the same way and rings restated from
    obstacles.py:269-281 (float64, unrounded)."""
    w, life = scenario
    run = HeadlessRunner(sfm_config, w, life, device_vehicles=True)
    centres = [w.veh_center.copy()]

    def vehicles_at(step):
        while len(centres) <= step:
            centres.append(centres[-1] + w.veh_vel * w.step_length)
        c = centres[step]
        rings = [LO.ellipse_ring(c[v], w.veh_yaw[v], w.veh_extent[v, 0], w.veh_extent[v, 1], w.veh_resolution)
                 for v in range(len(c))]
        return (None, list(c), list(w.veh_yaw), list(w.veh_vel), list(w.veh_extent), rings)

    steps = 60
    scene = O.Scene(sfm_config, w.step_length, w.borders, w.section_center, w.section_length, w.static_obstacles)
    want = LO.run_headless(scene, w, life, steps, vehicles_at=vehicles_at)
    for k in range(steps):
        run.tick()
        s = run.snapshot()
        assert np.array_equal(s['mode'], want['mode'][k + 1]), f'modes differ after tick {k}'
        assert np.array_equal(s['wp'], want['wp'][k + 1])
    assert np.abs(s['loc'] - want['loc'][-1]).max() < 1e-2
    got_c, _ = run.ctx.download_vehicles()
    assert np.array_equal(got_c, centres[steps - 1])


def test_engine_tick_with_device_vehicles_equals_headless_runner(scenario, sfm_config):
    """Engine.tick / Engine.step with device-resident vehicles use ONE vehicle state per tick for the gap acceptance and
    the dynamic-obstacle force (run_simulation.py:92-102), tick 0 seeing the uploaded state: bit-identical to the
    HeadlessRunner sequence (which the test above pins to the oracle), tick by tick."""
    from sfm_b200.engine import Engine
    w, life = scenario
    run = HeadlessRunner(sfm_config, w, life, device_vehicles=True)
    eng = Engine(sfm_config, w.step_length, device=0)
    eng.load(w, device_vehicles=True)
    eng.load_lifecycle(w, life)
    # (different staging origins -- (0, 0) vs the crowd centre -- on purpose: the (hi, lo) staging is origin-independent)
    for k in range(40):
        run.tick()
        eng.tick()
        s = run.snapshot()
        loc, vel = eng.local_state()
        assert np.array_equal(eng.ctx.download_mode_codes(), s['mode']), f'modes differ after tick {k}'
        assert np.array_equal(loc, s['loc']) and np.array_equal(vel, s['vel']), f'state differs after tick {k}'
        assert np.array_equal(eng.ctx.download_vehicles()[0], run.ctx.download_vehicles()[0])
    # Engine.step (no mode machines) keeps the same vehicle clock
    eng2 = Engine(sfm_config, w.step_length, device=0)
    eng2.load(w, device_vehicles=True)
    eng2.step(3, True)
    want = w.veh_center.copy()
    for _ in range(2):
        want = want + w.veh_vel * w.step_length
    assert np.array_equal(eng2.ctx.download_vehicles()[0], want)       # 3 ticks = states 0, 1, 2


def test_despawn_on_arrival_matches_reference_golden(scenario, sfm_config):
    """sfm_despawn_finished: the device removes the pedestrians the reference destroys, at the same ticks, keeps the row
    order, and the shrinking crowd keeps producing the reference's modes."""
    import dataclasses
    w, life = scenario
    g = np.load(os.path.join(GOLDEN, 'lifecycle_despawn.npz'))
    run = HeadlessRunner(sfm_config, w, dataclasses.replace(life, despawn_on_arrival=True))
    for k in range(LIFECYCLE_STEPS):
        run.tick()
        alive = np.nonzero(g['alive'][k + 1])[0]
        assert np.array_equal(run.ids, alive), f'crowd differs after tick {k}'
        assert run.ctx.n == len(alive)
        assert np.array_equal(run.ctx.download_mode_codes(), g['mode'][k + 1][alive]), f'modes differ after tick {k}'
    s = run.snapshot()
    assert np.array_equal(run.ids, g['ids_final']) and np.array_equal(s['wp'], g['wp_final'])
    assert np.abs(s['loc'] - g['loc_final']).max() < 0.1 and np.median(np.abs(s['loc'] - g['loc_final'])) < 1e-4


def test_spawn_and_despawn_match_reference_golden(sfm_config):
    """sfm_append_pedestrians + sfm_despawn_finished in one run: pedestrians join in two waves and leave on arrival; the
    device crowd has the reference's members, in the reference's row order, with the reference's modes at every tick."""
    import dataclasses
    w, life = synth.make_lifecycle(spawn_late=14)
    life = dataclasses.replace(life, despawn_on_arrival=True)
    g = np.load(os.path.join(GOLDEN, 'lifecycle_spawn.npz'))
    run = HeadlessRunner(sfm_config, w, life)
    assert run.ctx.n == 34
    for k in range(100):
        run.tick()
        alive = np.nonzero(g['alive'][k + 1])[0]
        assert sorted(run.ids) == alive.tolist() and run.ctx.n == len(alive), f'crowd differs after tick {k}'
        assert np.array_equal(run.ctx.download_mode_codes(), g['mode'][k + 1][run.ids]), f'modes differ after tick {k}'
    assert np.array_equal(run.ids, g['ids_final'])
    s = run.snapshot()
    assert np.array_equal(s['wp'], g['wp_final']) and np.abs(s['loc'] - g['loc_final']).max() < 0.1


def test_headless_run_writes_reference_csv_files(scenario, sfm_config, tmp_path):
    """A recorded device-resident run ends in the reference's four CSV files (output_generator.py:32-110)."""
    w, life = scenario
    run = HeadlessRunner(sfm_config, w, life, record_every=5, record_capacity=8)
    first = None
    for k in range(40):
        if k == 5:
            loc5, vel5 = run.ctx.download_state()
        run.tick()
        if k == 5:
            first = (loc5, vel5, run.ctx.download_frames()[1][1])
    out = run.write_csv(str(tmp_path), 'headless')
    ped = open(os.path.join(out, 'pedestrian.csv')).read().splitlines()
    veh = open(os.path.join(out, 'vehicle.csv')).read().splitlines()
    assert ped[0] == 'ped_id,frame,time,x,y,v_x,v_y,mode' and len(ped) == 1 + 8 * w.n
    assert veh[0] == 'veh_id,frame,time,x,y,heading,vel,ext_x,ext_y' and len(veh) == 1 + 8 * len(w.veh_center)
    # frame 1 was recorded at tick 5, after the machines ticked and before the forces moved anybody (:75-81)
    row = ped[1 + w.n + 3].split(',')
    assert row[:2] == ['3', '1'] and float(row[2]) == 0.25
    assert [float(v) for v in row[3:7]] == [first[0][3, 0], first[0][3, 1], first[1][3, 0], first[1][3, 1]]
    assert np.array_equal(first[2][:, :2], first[0][:, :2])
    assert len(open(os.path.join(out, 'borders.csv')).read().splitlines()) == 1 + sum(len(b) for b in w.borders)
    assert len(open(os.path.join(out, 'obstacles.csv')).read().splitlines()) == 1 + sum(len(r) for _, r in w.static_obstacles)


def test_recording_across_spawn_and_despawn(sfm_config, tmp_path):
    """The reference records a full snapshot every tick whatever happens to the crowd (pedestrian_state.py:100-104): frames
    recorded before a spawn / despawn survive it, every frame carries the ids of the crowd it shows, and pedestrian.csv
    has one row per pedestrian per frame with running frame numbers."""
    import dataclasses
    w, life = synth.make_lifecycle(spawn_late=14)
    life = dataclasses.replace(life, despawn_on_arrival=True)
    ticks = 100
    run = HeadlessRunner(sfm_config, w, life, record_every=1, record_capacity=ticks)
    twin = HeadlessRunner(sfm_config, w, life)                     # same run, no recorder: supplies the expected frames
    want = []
    for k in range(ticks):
        before, ids = twin.ctx.download_state(), twin.ids.copy()
        late = np.nonzero(twin.spawn_tick == k)[0] if k > 0 else np.zeros(0, dtype=np.int64)
        want.append((np.concatenate((ids, late)), np.concatenate((before[0][:, :2], w.loc[late][:, :2])),
                     np.concatenate((before[1][:, :2], w.vel[late][:, :2]))))
        twin.tick()
        run.tick()
    segments = run.recorded_frames()
    assert len(segments) > 3                                       # the crowd changed several times
    frames = [(ids, xyv[f]) for times, xyv, mode, ids in segments for f in range(len(times))]
    times = np.concatenate([s[0] for s in segments])
    assert len(frames) == ticks and np.allclose(times, np.arange(ticks) * w.step_length)
    for k, ((ids, xyv), (want_ids, want_xy, want_v)) in enumerate(zip(frames, want)):
        assert np.array_equal(ids, want_ids), f'frame {k}: crowd differs'
        assert np.array_equal(xyv[:, :2], want_xy) and np.array_equal(xyv[:, 2:], want_v), f'frame {k}: state differs'
    assert np.array_equal(run.ids, twin.ids)
    out = run.write_csv(str(tmp_path), 'spawn_despawn')
    ped = open(os.path.join(out, 'pedestrian.csv')).read().splitlines()
    assert len(ped) == 1 + sum(len(ids) for ids, _ in frames)
    last = ped[-1].split(',')
    assert int(last[1]) == ticks - 1 and int(last[0]) == frames[-1][0][-1]


def test_degenerate_gap_acceptance_on_device(sfm_config):
    """K4a on the collinear / zero-length scenes of tests/degenerate_traffic.py: the decisions of the oracle (shapely's
    overlap-segment semantics, check_traffic.py:46-58)."""
    from tests.degenerate_traffic import scenes
    ctx = native.Context(0)
    ctx.set_params(native.params_from_config(sfm_config, 0.05))
    for k, (loc, goal, speed, margin, centres, vels, extents, want) in enumerate(scenes()):
        expect = LO.check_traffic(loc, goal, speed, margin, centres, vels, extents)
        assert want is None or expect == want
        n = 3                                                  # the scene's pedestrian between two bystanders
        loc3, wp3 = np.zeros((n, 3)), np.zeros((n, 3))
        loc3[:, :2], wp3[:, :2] = [[-50.0, 40.0], loc, [60.0, -40.0]], [[-50.0, 45.0], goal, [60.0, -45.0]]
        mode = np.array([1, 4, 1], dtype=np.uint8)
        ctx.upload_state(loc3, np.zeros((n, 3)), wp3, np.full(n, 0.3), np.full(n, speed / 1.5), mode)
        ctx.set_mode_machines(np.full(n, speed / 1.5), np.full(n, speed), np.full(n, margin))
        ctx.set_traffic(centres, vels, np.tile(extents[0], (len(centres), 1)))
        ctx.tick_modes(0.0)
        got = ctx.download_mode_codes()
        assert (got[1] == 2) == expect, f'scene {k + 1}: device {got[1]}, oracle says cross={expect}'
        assert got[0] == 1 and got[2] == 1
