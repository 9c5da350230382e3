"""Pins the float64 oracle against the imported, unmodified reference (only where /root/reference exists).

On the GPU box these tests skip; the same comparison is then carried by the committed golden vectors
(tests/test_oracle_golden.py), which oracle/make_golden.py produced from the imported reference.
"""
import numpy as np
import pytest

from oracle import ref_loader, sfm_oracle as O
from sfm_b200 import synth

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason='reference tree not present')


@pytest.fixture(scope='module')
def ref():
    return ref_loader.load()


def _oracle_forces(w, cfg, step=0):
    scene = O.Scene(cfg, w.step_length, w.borders, w.section_center, w.section_length, w.static_obstacles)
    veh = w.vehicles_at(step)
    dyn = list(zip(veh[1], veh[5])) if veh else None
    return scene, O.forces_by_class(scene, w.loc, w.vel, w.next_waypoint, w.radius, w.target_speed, w.mode, dyn,
                                    veh[3] if veh else None)


@pytest.mark.parametrize('use_radius', [False, True])
@pytest.mark.parametrize('z_spread', [0.0, 0.2])
def test_force_classes_match_reference(ref, sfm_config, use_radius, z_spread):
    cfg = dict(sfm_config, use_ped_radius=use_radius)
    w = synth.make_config(2, n=512, z_spread=z_spread)
    sim = ref_loader.build_simulation(ref, w, cfg)
    sim.update_dynamic_obstacles(w.vehicles_at(0))
    _, mine = _oracle_forces(w, cfg)
    assert list(mine) == list(sim.forces)                       # dict order accel, ped, border, static, dynamic
    for name, force in sim.forces.items():
        want = force.get_force(sim.peds)
        np.testing.assert_allclose(mine[name], want, rtol=1e-12, atol=1e-12, err_msg=name)


def test_cfg1_trajectory_matches_reference(ref, sfm_config):
    w = synth.make_config(1)
    got = ref_loader.run_ticks(ref, w, sfm_config, 100, record_forces=False)
    scene = O.Scene(sfm_config, w.step_length, w.borders, w.section_center, w.section_length, w.static_obstacles)
    loc, vel = w.loc.copy(), w.vel.copy()
    for step in range(100):
        veh = w.vehicles_at(step)
        loc, vel, _ = O.step(scene, loc, vel, w.next_waypoint, w.radius, w.target_speed, w.mode,
                             list(zip(veh[1], veh[5])), veh[3])
        np.testing.assert_allclose(vel, got['vel'][step + 1], rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(loc, got['loc'][step + 1], rtol=1e-9, atol=1e-10)


def test_force_switches(ref, sfm_config):
    cfg = dict(sfm_config, forces=dict(acceleration_force=True, pedestrian_force=False, border_force=True,
                                       static_obstacle_force=False, dynamic_obstacle_force=False))
    w = synth.make_config(1)
    sim = ref_loader.build_simulation(ref, w, cfg)
    _, mine = _oracle_forces(w, cfg)
    assert list(mine) == list(sim.forces) == ['acceleration_force', 'border_force']


@pytest.mark.parametrize('use_radius', [False, True])
def test_appendix_b_edge_cases_match_reference(ref, sfm_config, use_radius):
    """SURVEY.md appendix B side by side (tests/appendix_b.py): argmin ties, strict cutoffs at exactly the cutoff distance,
    theta = -pi / +pi without a wrap, zero distances to a border / ring point, masked modes, waypoint reached, zero target
    speed, |v'| = 0 -- every class and one tick of new velocities, oracle == imported reference."""
    from tests import appendix_b
    cfg = dict(sfm_config, use_ped_radius=use_radius)
    w = appendix_b.scene()
    sim = ref_loader.build_simulation(ref, w, cfg)
    sim.update_dynamic_obstacles(w.vehicles_at(0))
    scene, mine = _oracle_forces(w, cfg)
    with np.errstate(all='ignore'):
        for name, force in sim.forces.items():
            want = force.get_force(sim.peds)
            np.testing.assert_allclose(mine[name], want, rtol=1e-12, atol=1e-12, err_msg=name)
        sim.tick(0.0)
    veh = w.vehicles_at(0)
    _, vel, _ = O.step(scene, w.loc, w.vel, w.next_waypoint, w.radius, w.target_speed, w.mode, list(zip(veh[1], veh[5])),
                       veh[3])
    np.testing.assert_allclose(vel, sim.get_new_velocities()['vel'], rtol=1e-12, atol=1e-12)
    # the cases are what the scene's docstring says they are
    b, s = mine['border_force'], mine['static_obstacle_force']
    assert not b[[5, 8, 9]].any() and not s[5].any() and not vel[1].any() and not vel[2].any()
    _, pairs_b = O.border_force(w.loc, w.radius, w.mode, w.borders, w.section_center, w.section_length, scene.border,
                                use_radius, return_pairs=True)
    _, pairs_s = O.obstacle_force(w.loc, w.vel, w.radius, [c for c, _ in w.static_obstacles],
                                  [r for _, r in w.static_obstacles], None, scene.static, use_radius, return_pairs=True)
    pb = {(int(p), int(k)): int(q) for p, k, q in pairs_b}
    ps = {(int(p), int(k)): int(q) for p, k, q in pairs_s}
    assert pb[(3, 0)] == 4 and ps[(4, 0)] == 0                     # first index on exact ties
    assert (10, 1) not in pb and (11, 1) not in ps                 # exactly at the cutoff: excluded
    assert pb[(12, 0)] == 7 and ps[(13, 0)] == 2                   # zero distance
