"""Lifecycle oracle (mode machines, gap acceptance, waypoint hand-over, vehicle rings, CSV) against the reference.

CPU only.  The golden file was produced by the reference's own classes (oracle/make_golden.py --only lifecycle); when
/root/reference is present the same run is repeated live.
"""
import os

import numpy as np
import pytest

from tests.golden_util import GOLDEN
from oracle import lifecycle_oracle as LO
from oracle import ref_loader
from oracle import sfm_oracle as O
from oracle.make_golden import LIFECYCLE_STEPS, lifecycle_digest
from sfm_b200 import synth


@pytest.fixture(scope='module')
def scenario():
    return synth.make_lifecycle()


@pytest.fixture(scope='module')
def oracle_run(scenario, sfm_config):
    w, life = scenario
    scene = O.Scene(sfm_config, w.step_length, w.borders, w.section_center, w.section_length, w.static_obstacles)
    return LO.run_headless(scene, w, life, LIFECYCLE_STEPS)


def test_oracle_reproduces_reference_golden(scenario, oracle_run):
    w, life = scenario
    g = np.load(os.path.join(GOLDEN, 'lifecycle.npz'))
    assert str(g['digest']) == lifecycle_digest(w, life), 'synthetic scenario drifted from the golden inputs'
    assert np.array_equal(g['mode'], oracle_run['mode'])                  # every mode of every pedestrian at every tick
    assert np.array_equal(g['wp'], oracle_run['wp'])                      # hand-overs happen at the same ticks
    assert np.array_equal(g['target_speed'], oracle_run['target_speed'])
    assert np.array_equal(g['mode_speed'], oracle_run['mode_speed'])
    remaining = np.array([len(r) for r in life.routes])[None, :] - oracle_run['cursor']
    assert np.array_equal(g['remaining'], remaining)
    assert np.abs(g['loc'] - oracle_run['loc']).max() <= 1e-9 and np.abs(g['vel'] - oracle_run['vel']).max() <= 1e-9
    # the scenario exercises every transition of the machine (ped_mode_manager.py:37-69)
    m = g['mode'].astype(int)
    seen = {(a, b) for a, b in zip(m[:-1].ravel(), m[1:].ravel()) if a != b}
    assert {(0, 1), (1, 4), (4, 2), (2, 3), (3, 1)} <= seen
    assert ((m[:-1] == 4) & (m[1:] == 4)).sum() > 20                      # gap acceptance said "wait" many times


@pytest.mark.skipif(not ref_loader.available(), reason='reference tree not present')
def test_oracle_matches_imported_reference_live(scenario, oracle_run, sfm_config):
    w, life = scenario
    ref = ref_loader.load()
    r = ref_loader.run_lifecycle(ref, w, life, ref_loader.load_config(), 60)
    for key in ('mode', 'wp', 'mode_speed'):
        assert np.array_equal(r[key], oracle_run[key][:61]), key
    assert np.array_equal(r['target_speed'], oracle_run['target_speed'][:60])
    assert np.abs(r['loc'] - oracle_run['loc'][:61]).max() <= 1e-9


@pytest.mark.skipif(not ref_loader.available(), reason='reference tree not present')
def test_mode_machine_columns_equal_reference_objects():
    """Random request sequences through the reference's PedModeManager objects and through the array restatement."""
    ref = ref_loader.load()
    PedMode, Manager = ref.ped_mode_manager.PedMode, ref.ped_mode_manager.PedModeManager
    rng = np.random.default_rng(5)
    n = 64
    speed, factor, margin = rng.uniform(1, 2, n), rng.uniform(1, 2, n), rng.uniform(-1, 2, n)
    init = rng.integers(0, 5, n)
    objs = [Manager(f'p_{i}', speed[i], PedMode(int(init[i])), factor[i], margin[i]) for i in range(n)]
    cols = LO.Machines.create(speed, init, factor, margin)
    t = 0.0
    for _ in range(200):
        t += rng.uniform(0.0, 1.5)
        for o in objs:
            o.tick(t)
        cols.tick(t)
        rows = rng.choice(n, size=8, replace=False)
        wanted = int(rng.integers(0, 5))
        for i in rows:
            objs[i].set_mode(PedMode(wanted))
        cols.set_mode(rows, wanted)
        assert [int(o.current_mode) for o in objs] == cols.mode.tolist()
        assert [float(o.target_speed) for o in objs] == cols.target_speed.tolist()
        assert [float(o.next_mode_time) for o in objs] == cols.next_mode_time.tolist()


def test_host_check_traffic_equals_oracle():
    """The drop-in's check_traffic.py (host, shapely-free) and the oracle agree on random scenes."""
    import check_traffic as host
    from ped_mode_manager import PedMode, PedModeManager
    rng = np.random.default_rng(11)
    dtype = [('loc', 'f8', (3,)), ('next_waypoint', 'f8', (3,)), ('mode', 'O')]
    decisions = []
    for _ in range(300):
        v = int(rng.integers(1, 6))
        centres, vels = rng.uniform(0, 30, (v, 2)), rng.normal(0, 6, (v, 2))
        vels[rng.random(v) < 0.2] = 0.0
        extents = np.tile([2.4, 1.0], (v, 1))
        ped = np.zeros(1, dtype=dtype)[0]
        ped['loc'][:2], ped['next_waypoint'][:2] = rng.uniform(0, 30, 2), rng.uniform(0, 30, 2)
        ped['mode'] = PedModeManager('p_0', 1.3, PedMode.CHECKING_TRAFFIC, 1.5, float(rng.uniform(-0.5, 2.5)))
        want = LO.check_traffic(ped['loc'], ped['next_waypoint'], ped['mode'].crossing_speed,
                                ped['mode'].crossing_safety_margin, centres, vels, extents)
        got = host.check_traffic(ped, [(c, None) for c in centres], list(vels), list(extents))
        assert got == want
        decisions.append(want)
    assert 0.05 < np.mean(decisions) < 0.95


def test_ellipse_ring_restatement():
    """obstacles.py:269-281: point count rule and geometry (synth.ellipse_ring is the vectorised host version)."""
    for centre, yaw, ex, ey, res in (((3.0, -2.0), 30.0, 2.4, 1.0, 0.1), ((0.0, 0.0), -135.0, 0.2, 0.1, 0.5),
                                     ((10.0, 5.0), 0.0, 2.4, 1.0, 0.17)):
        ring = LO.ellipse_ring(centre, yaw, ex, ey, res)
        assert len(ring) == max(6, int((2 * ex + 2 * ey) / res))
        assert np.abs(ring - synth.ellipse_ring(centre, yaw, ex, ey, res)).max() < 1e-12
        # points lie on the rotated ellipse with semi-axes sqrt(2) * extent
        c, s = np.cos(np.radians(yaw)), np.sin(np.radians(yaw))
        local = (ring - centre) @ np.array([[c, -s], [s, c]])
        assert np.allclose((local[:, 0] / (ex * np.sqrt(2))) ** 2 + (local[:, 1] / (ey * np.sqrt(2))) ** 2, 1.0)


def test_output_generator_bytes_equal_reference(tmp_path):
    """The drop-in OutputGenerator writes byte-identical files to the reference's (golden: its own run, build container)."""
    import types
    from output_generator import OutputGenerator
    scene = synth.make_output_scene()
    ped_sim = types.SimpleNamespace(peds=types.SimpleNamespace(all_states=scene['ped_states']),
                                    all_dyn_obs_states=scene['veh_states'], static_obstacles=scene['static_obstacles'],
                                    borders=scene['borders'])
    gen = OutputGenerator(ped_sim, str(tmp_path), 'golden')
    gen.generate_ped_csv(); gen.generate_veh_csv(); gen.generate_borders_csv(); gen.generate_obstacles_csv()
    g = np.load(os.path.join(GOLDEN, 'output_csv.npz'))
    for name in ('pedestrian.csv', 'vehicle.csv', 'borders.csv', 'obstacles.csv'):
        with open(os.path.join(gen.output_dir, name), 'rb') as f:
            assert f.read() == g[name.replace('.', '_')].tobytes(), name
    # device-frame path: same bytes from (times, xyv, mode) arrays
    states = scene['ped_states']
    times = np.array(list(states))
    xyv = np.array([np.column_stack((s['loc'][:, :2], s['vel'][:, :2])) for s in states.values()])
    mode = np.array([[int(m) for m in s['mode']] for s in states.values()], dtype=np.uint8)
    ids = [int(nm.split('_')[-1]) for nm in next(iter(states.values()))['name']]
    gen2 = OutputGenerator(ped_sim, str(tmp_path / 'dev'), 'golden')
    gen2.generate_ped_csv(device_frames=(times, xyv, mode), ped_ids=ids)
    with open(os.path.join(gen2.output_dir, 'pedestrian.csv'), 'rb') as f:
        assert f.read() == g['pedestrian_csv'].tobytes()


def test_oracle_despawn_reproduces_reference_golden(scenario, sfm_config):
    """despawn_on_arrival: the oracle removes the same pedestrians at the same ticks as the reference's
    destroy_pedestrian calls and the survivors end in the same state."""
    from oracle.make_golden import pack_ragged
    w, life = scenario
    scene = O.Scene(sfm_config, w.step_length, w.borders, w.section_center, w.section_length, w.static_obstacles)
    got = pack_ragged(LO.run_headless(scene, w, life, LIFECYCLE_STEPS, despawn=True), w.n)
    g = np.load(os.path.join(GOLDEN, 'lifecycle_despawn.npz'))
    assert str(g['digest']) == lifecycle_digest(w, life)
    assert np.array_equal(g['alive'], got['alive']) and np.array_equal(g['mode'], got['mode'])
    assert np.array_equal(g['ids_final'], got['ids_final']) and 0 < len(got['ids_final']) < w.n // 2
    assert np.abs(g['loc_final'] - got['loc_final']).max() <= 1e-9 and np.array_equal(g['wp_final'], got['wp_final'])


def test_host_mode_manager_equals_oracle_columns():
    """The drop-in's table-driven PedModeManager against the array restatement (itself pinned to the reference's
    objects above) under random tick / request sequences, unknown modes included."""
    from ped_mode_manager import PedMode, PedModeManager
    rng = np.random.default_rng(23)
    n = 48
    speed, factor, margin = rng.uniform(1, 2, n), rng.uniform(1, 2, n), rng.uniform(-1, 2, n)
    init = rng.integers(0, 5, n)
    objs = [PedModeManager(f'p_{i}', speed[i], PedMode(int(init[i])), factor[i], margin[i]) for i in range(n)]
    cols = LO.Machines.create(speed, init, factor, margin)
    t = 0.0
    for _ in range(300):
        t += rng.uniform(0.0, 1.5)
        for o in objs:
            o.tick(t)
        cols.tick(t)
        rows = rng.choice(n, size=6, replace=False)
        wanted = int(rng.integers(0, 6))                       # 5 is not a mode: ignored by both
        for i in rows:
            objs[i].set_mode(PedMode(wanted) if wanted < 5 else wanted)
        cols.set_mode(rows, wanted)
        assert [int(o.current_mode) for o in objs] == cols.mode.tolist()
        assert [float(o.target_speed) for o in objs] == cols.target_speed.tolist()
        assert [float(o.next_mode_time) for o in objs] == cols.next_mode_time.tolist()


def test_oracle_spawn_and_despawn_reproduce_reference_golden(sfm_config):
    """Two late spawn waves plus despawn on arrival: same crowd membership, modes and final state as the reference."""
    import dataclasses
    from oracle.make_golden import pack_ragged
    w, life = synth.make_lifecycle(spawn_late=14)
    life = dataclasses.replace(life, despawn_on_arrival=True)
    scene = O.Scene(sfm_config, w.step_length, w.borders, w.section_center, w.section_length, w.static_obstacles)
    got = pack_ragged(LO.run_headless(scene, w, life, 100, despawn=True), w.n)
    g = np.load(os.path.join(GOLDEN, 'lifecycle_spawn.npz'))
    assert str(g['digest']) == lifecycle_digest(w, life) and np.array_equal(g['spawn_tick'], life.spawn_tick)
    assert np.array_equal(g['alive'], got['alive']) and np.array_equal(g['mode'], got['mode'])
    assert g['alive'][0].sum() == 34 and g['alive'][46].sum() > g['alive'][44].sum()        # the second wave arrives
    assert np.array_equal(g['ids_final'], got['ids_final']) and np.abs(g['loc_final'] - got['loc_final']).max() <= 1e-9


def test_degenerate_gap_acceptance_scenes():
    ref = ref_loader.load() if ref_loader.available() else None
    _degenerate_scenes(ref)


def _degenerate_scenes(ref):
    """Collinear and zero-length paths (tests/degenerate_traffic.py): the oracle, the drop-in's host function and -- where
    the reference tree is present -- the reference's own check_traffic (over the LineString stand-in, which returns the
    overlap segment like shapely) take the hand-derived decisions."""
    import check_traffic as host
    from ped_mode_manager import PedMode, PedModeManager
    from tests.degenerate_traffic import scenes
    dtype = [('loc', 'f8', (3,)), ('next_waypoint', 'f8', (3,)), ('mode', 'O')]
    for k, (loc, goal, speed, margin, centres, vels, extents, want) in enumerate(scenes()):
        got = LO.check_traffic(loc, goal, speed, margin, centres, vels, extents)
        if want is not None:
            assert got == want, f'scene {k + 1}'
        ped = np.zeros(1, dtype=dtype)[0]
        ped['loc'][:2], ped['next_waypoint'][:2] = loc, goal
        ped['mode'] = PedModeManager('p_0', speed / 1.5, PedMode.CHECKING_TRAFFIC, 1.5, margin)
        vehicles = [(c, None) for c in centres]
        assert host.check_traffic(ped, vehicles, list(vels), list(extents)) == got, f'host, scene {k + 1}'
        if ref is not None:
            import importlib.util
            import os
            spec = importlib.util.spec_from_file_location('_ref_check_traffic',
                                                          os.path.join(ref_loader.REFERENCE_DIR, 'check_traffic.py'))
            mod = importlib.util.module_from_spec(spec)
            import sys
            saved = sys.modules.get('stateutils')
            sys.modules['stateutils'] = ref.stateutils
            try:
                spec.loader.exec_module(mod)
            finally:
                if saved is not None:
                    sys.modules['stateutils'] = saved
                else:
                    sys.modules.pop('stateutils', None)
            rped = np.zeros(1, dtype=dtype)[0]
            rped['loc'][:2], rped['next_waypoint'][:2] = loc, goal
            rped['mode'] = ref.ped_mode_manager.PedModeManager('p_0', speed / 1.5, ref.ped_mode_manager.PedMode.CHECKING_TRAFFIC,
                                                               1.5, margin)
            assert mod.check_traffic(rped, vehicles, vels, extents) == got, f'reference, scene {k + 1}'
