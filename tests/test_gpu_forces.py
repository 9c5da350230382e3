"""GPU parity of every force class against the float64 oracle and the reference's golden vectors (through the C ABI).

Tolerances: the float32 all-pairs kernel must meet BASELINE.json's 1e-4 relative / 1e-5 m/s^2 absolute per component;
the float64 cell-list kernels and the integrate kernel are held to 1e-11 (they follow numpy's operation order).
"""
import numpy as np
import pytest

from oracle import sfm_oracle as O
from sfm_b200 import native, synth
from tests import golden_util as G
from tests.gpu_util import assert_forces_close, make_context

pytestmark = pytest.mark.gpu


def _oracle_classes(w, cfg, step=0, rows=None):
    scene = G.scene_for(w, cfg)
    dyn, dyn_vel = G.dyn_for(w, step)
    return O.forces_by_class(scene, w.loc, w.vel, w.next_waypoint, w.radius, w.target_speed, w.mode, dyn, dyn_vel,
                             rows=rows)


@pytest.mark.parametrize('use_radius', [False, True])
def test_cfg1_all_classes(sfm_config, use_radius):
    cfg = dict(sfm_config, use_ped_radius=use_radius)
    w = synth.make_config(1)
    ctx = make_context(w, cfg)
    want = _oracle_classes(w, cfg)
    _, risk = O.pedestrian_force(w.loc, w.vel, w.radius, G.scene_for(w, cfg).ped, use_radius, return_risk=True)
    for k, name in enumerate(native.FORCE_CLASSES):
        got = ctx.force(k)
        if name == 'pedestrian_force':
            assert_forces_close(got, want[name], risk=risk, name=name)
        else:
            np.testing.assert_allclose(got, want[name], rtol=1e-11, atol=1e-11, err_msg=name)


@pytest.mark.parametrize('use_radius,z', [(False, 0), (True, 0), (False, 1), (True, 1)])
def test_cfg2_against_reference_golden(sfm_config, use_radius, z):
    """N = 4096, all five classes, vs arrays the imported reference produced (tests/golden)."""
    cfg = dict(sfm_config, use_ped_radius=use_radius)
    w = synth.make_config(2, z_spread=0.2 if z else 0.0)
    g = G.load(f'cfg2_forces_r{int(use_radius)}_z{z}.npz', w)
    ctx = make_context(w, cfg)
    rows = np.arange(0, w.n, 4)
    _, risk = O.pedestrian_force(w.loc, w.vel, w.radius, G.scene_for(w, cfg).ped, use_radius, rows=rows,
                                 return_risk=True)
    for k, name in enumerate(native.FORCE_CLASSES):
        got, want = ctx.force(k), g[f'F_{name}']
        if name == 'pedestrian_force':
            assert np.isfinite(got).all()                   # unsampled rows are only checked for finiteness
            assert_forces_close(got[rows], want[rows], risk=risk, name=name)
        else:
            np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-11, err_msg=name)


@pytest.mark.parametrize('cls,name', [(native.BORDER, 'border'), (native.STATIC_OBSTACLE, 'static'),
                                      (native.DYNAMIC_OBSTACLE, 'dynamic')])
def test_neighbour_enumeration_bit_exact(sfm_config, cls, name):
    """The cell-list kernels must visit exactly the reference's (pedestrian, item, nearest point) triplets."""
    w = synth.make_config(2)
    ctx = make_context(w, sfm_config)
    scene = G.scene_for(w, sfm_config)
    if cls == native.BORDER:
        _, want = O.border_force(w.loc, w.radius, w.mode, scene.borders, scene.section_center, scene.section_length,
                                 scene.border, False, return_pairs=True)
    elif cls == native.STATIC_OBSTACLE:
        _, want = O.obstacle_force(w.loc, w.vel, w.radius, [c for c, _ in w.static_obstacles],
                                   [r for _, r in w.static_obstacles], None, scene.static, False, return_pairs=True)
    else:
        dyn, dyn_vel = G.dyn_for(w, 0)
        _, want = O.obstacle_force(w.loc, w.vel, w.radius, [c for c, _ in dyn], [r for _, r in dyn], dyn_vel,
                                   scene.dynamic, False, return_pairs=True)
    got = ctx.enumerate_pairs(cls)
    assert len(want) > 1000
    np.testing.assert_array_equal(got, want)


def test_force_switches_and_empty_sets(sfm_config):
    w = synth.make_config(1)
    w.borders, w.static_obstacles, w.veh_center = [], [], None
    ctx = make_context(w, sfm_config)
    for cls in (native.BORDER, native.STATIC_OBSTACLE, native.DYNAMIC_OBSTACLE):
        assert not ctx.force(cls).any()                      # forces.py:140-141, :209-210


def test_coincident_and_degenerate_pairs(sfm_config):
    """Zero distance, equal velocities, |D| = 0 -- same values as numpy wherever numpy is finite."""
    loc = np.array([[0, 0, 0], [0, 0, 0], [2, 0, 0], [2, 0, 0.5], [5, 5, 0], [7, 5, 0]], dtype=np.float64)
    vel = np.array([[1, 0, 0], [0, 1, 0], [0.5, 0, 0], [0.5, 0, 0], [0, 0, 0], [0.5, 0, 0]], dtype=np.float64)
    n = len(loc)
    w = synth.make_config(1)
    w.loc, w.vel = loc, vel
    w.next_waypoint, w.radius = np.zeros((n, 3)), np.full(n, 0.25)
    w.target_speed, w.mode = np.full(n, 1.3), np.ones(n, dtype=np.uint8)
    ctx = make_context(w, sfm_config)
    with np.errstate(all='ignore'):
        want = O.pedestrian_force(loc, vel, w.radius, G.scene_for(w, sfm_config).ped, False)
    got = ctx.force(native.PEDESTRIAN)
    finite = np.isfinite(want).all(axis=1)
    assert finite.sum() >= 4
    assert_forces_close(got[finite], want[finite], name='degenerate')


def test_degenerate_pairs_across_tiles(sfm_config):
    """Coincident pedestrians that live in different 256-row tiles, a vertical-only offset and a |D| = 0 pair: the
    unguarded symmetric fast path must hand exactly those rows to the repair path (sfm_stats.fixup_rows) and every row
    must match numpy wherever numpy is finite."""
    w = synth.make_config(2, n=1024)
    loc, vel = w.loc.copy(), w.vel.copy()
    loc[700] = loc[3]                                  # |d| = 0 across tiles 0 and 2
    loc[900, :2] = loc[20, :2]
    loc[900, 2] = 0.5                                  # d_xy = 0, d_z != 0  (angle_xy(e) = atan2(0, 0) = 0)
    loc[400] = loc[130] + np.array([2.0, 0.0, 0.0])
    vel[130], vel[400] = np.array([0.0, 0.0, 0.0]), np.array([0.5, 0.0, 0.0])      # lambda (v_i - v_j) = -e: |D| = 0
    w.loc, w.vel = loc, vel
    ctx = make_context(w, sfm_config)
    ctx.reset_stats()
    got = ctx.force(native.PEDESTRIAN)
    repaired = ctx.stats()['fixup_rows']
    with np.errstate(all='ignore'):
        want, risk = O.pedestrian_force(loc, vel, w.radius, G.scene_for(w, sfm_config).ped, False, return_risk=True)
    finite = np.isfinite(want).all(axis=1)
    assert finite.sum() >= w.n - 2
    assert 4 <= repaired <= 8, repaired
    assert_forces_close(got[finite], want[finite], risk=risk[finite], name='degenerate-across-tiles')


def test_planar_and_general_paths_agree(sfm_config):
    """A flat crowd takes the z-free fast path; lifting one pedestrian by 0 m with a non-zero origin forces the general
    path on the same numbers -- both must give the same forces to float32 rounding."""
    w = synth.make_config(2, n=2048)
    flat = make_context(w, sfm_config)
    f_flat = flat.force(native.PEDESTRIAN)
    general = native.Context(0)
    general.set_params(native.params_from_config(sfm_config, w.step_length))
    general.set_origin(0.0, 0.0, -1.0)                # staged z = 1 everywhere: non-planar flag set, same geometry
    general.upload_state(w.loc, w.vel, w.next_waypoint, w.radius, w.target_speed, w.mode)
    f_gen = general.force(native.PEDESTRIAN)
    np.testing.assert_allclose(f_flat, f_gen, rtol=2e-5, atol=2e-6)
    assert np.abs(f_flat[:, 2]).max() == 0.0


def test_pair_force_is_reproducible(sfm_config):
    """Integer (fixed-point) accumulation: two independent contexts give bit-identical pair forces."""
    w = synth.make_config(2, n=3000)
    a = make_context(w, sfm_config).force(native.PEDESTRIAN)
    b = make_context(w, sfm_config).force(native.PEDESTRIAN)
    np.testing.assert_array_equal(a, b)


def test_large_dynamic_set_enumeration_bit_exact(sfm_config):
    """A dynamic set of 100,000 obstacles re-binned on the device by the many-CTA radix sort / scan: the neighbour
    enumeration equals the reference's triplet for triplet, the forces agree to 1e-10 (forces.py:208-283)."""
    rng = np.random.default_rng(77)
    n, n_obs, side = 2048, 100_000, 1000.0
    w = synth.make_config(2, n=n)
    loc = w.loc.copy()
    loc[:, :2] = rng.uniform(0.0, side, size=(n, 2))
    centres = rng.uniform(0.0, side, size=(n_obs, 2))
    t = 2.0 * np.pi * np.arange(6) / 6
    ring0 = 0.3 * np.column_stack((np.cos(t), np.sin(t)))
    rings = [c + ring0 for c in centres]
    vels = rng.normal(0.0, 3.0, size=(n_obs, 2))
    ctx = native.Context(0)
    ctx.set_params(native.params_from_config(sfm_config, w.step_length))
    ctx.upload_state(loc, w.vel, w.next_waypoint, w.radius, w.target_speed, w.mode)
    offsets = 6 * np.arange(n_obs + 1, dtype=np.int64)
    ctx.set_obstacles_csr(native.DYNAMIC_OBSTACLE, centres, vels, offsets, np.concatenate(rings))
    scene = G.scene_for(w, sfm_config)
    want_f, want = O.obstacle_force(loc, w.vel, w.radius, centres, rings, vels, scene.dynamic, False, return_pairs=True)
    got = ctx.enumerate_pairs(native.DYNAMIC_OBSTACLE, capacity=1 << 20)
    assert len(want) > 100_000
    np.testing.assert_array_equal(got, want)
    np.testing.assert_allclose(ctx.force(native.DYNAMIC_OBSTACLE), want_f, rtol=1e-10, atol=1e-10)
    # same answer when the set is uploaded again (re-sorted) and when a mid-sized set takes the many-CTA path
    ctx.set_obstacles_csr(native.DYNAMIC_OBSTACLE, centres, vels, offsets, np.concatenate(rings))
    np.testing.assert_array_equal(ctx.enumerate_pairs(native.DYNAMIC_OBSTACLE, capacity=1 << 20), want)


def test_many_cta_sort_and_scan_on_small_sets(sfm_config, monkeypatch):
    """SFM_SORT_SINGLE_MAX=0 sends every set through the many-CTA sort: the cfg2 enumerations stay bit-exact."""
    monkeypatch.setenv('SFM_SORT_SINGLE_MAX', '0')
    w = synth.make_config(2)
    ctx = make_context(w, sfm_config)
    scene = G.scene_for(w, sfm_config)
    _, want = O.border_force(w.loc, w.radius, w.mode, scene.borders, scene.section_center, scene.section_length,
                             scene.border, False, return_pairs=True)
    np.testing.assert_array_equal(ctx.enumerate_pairs(native.BORDER), want)
    _, want = O.obstacle_force(w.loc, w.vel, w.radius, [c for c, _ in w.static_obstacles],
                               [r for _, r in w.static_obstacles], None, scene.static, False, return_pairs=True)
    np.testing.assert_array_equal(ctx.enumerate_pairs(native.STATIC_OBSTACLE), want)


def test_perception_threshold_change_needs_a_new_upload(sfm_config):
    """The obstacle sets bake perception_threshold into their cutoffs and cell grid: after sfm_set_params changed it the
    resident set is refused (no silent use of the old cutoff) until it is uploaded again -- then it follows the new one."""
    w = synth.make_config(2, n=1024)
    ctx = make_context(w, sfm_config)
    before = ctx.force(native.STATIC_OBSTACLE)
    cfg = dict(sfm_config, static_obstacle_force=dict(sfm_config['static_obstacle_force'], perception_threshold=6.0))
    ctx.set_params(native.params_from_config(cfg, w.step_length))
    with pytest.raises(native.SfmError, match='perception_threshold changed'):
        ctx.force(native.STATIC_OBSTACLE)
    ctx.set_obstacles(native.STATIC_OBSTACLE, [c for c, _ in w.static_obstacles], [r for _, r in w.static_obstacles])
    scene = G.scene_for(w, cfg)
    want, pairs = O.obstacle_force(w.loc, w.vel, w.radius, [c for c, _ in w.static_obstacles],
                                   [r for _, r in w.static_obstacles], None, scene.static, False, return_pairs=True)
    np.testing.assert_array_equal(ctx.enumerate_pairs(native.STATIC_OBSTACLE), pairs)
    after = ctx.force(native.STATIC_OBSTACLE)
    np.testing.assert_allclose(after, want, rtol=1e-11, atol=1e-11)
    assert np.abs(after - before).max() > 1e-3                      # the smaller threshold really dropped neighbours
    ctx.set_params(native.params_from_config(sfm_config, w.step_length))        # back to the old threshold: stale again
    with pytest.raises(native.SfmError):
        ctx.step(1, True)


def test_infinite_section_length_reaches_every_pedestrian(sfm_config):
    """A section whose length (= cutoff, forces.py:149-150) is infinite is 'close' to every pedestrian of the crowd, not
    only to those in the neighbouring grid cells."""
    w = synth.make_config(2, n=2048)
    length = w.section_length.copy()
    length[3] = np.inf
    ctx = make_context(w, sfm_config)
    ctx.set_borders(w.borders, w.section_center, length)
    scene = G.scene_for(w, sfm_config)
    want, pairs = O.border_force(w.loc, w.radius, w.mode, w.borders, w.section_center, length, scene.border, False,
                                 return_pairs=True)
    assert (pairs[:, 1] == 3).sum() == w.n
    np.testing.assert_array_equal(ctx.enumerate_pairs(native.BORDER), pairs)
    np.testing.assert_allclose(ctx.force(native.BORDER), want, rtol=1e-11, atol=1e-11)


def test_cutoffs_are_strict_to_the_last_bit(sfm_config):
    """forces.py:149-150, :222-223 test `norm < cutoff` in float64: pedestrians at exactly the cutoff distance (also where
    sqrt rounds onto it), one ulp inside / outside, and just inside the float32 pre-filter's margin must be classified like
    numpy classifies them -- for a section, a static obstacle and a vehicle."""
    w = synth.make_config(2, n=256)
    centre = np.array([100.0, 200.0])
    cut_static = sfm_config['static_obstacle_force']['perception_threshold']            # 20
    cut_dynamic = sfm_config['dynamic_obstacle_force']['perception_threshold']          # 50
    ring = centre + 0.4 * np.column_stack((np.cos(np.arange(8) * np.pi / 4), np.sin(np.arange(8) * np.pi / 4)))
    line = centre + np.column_stack((np.linspace(-5, 5, 101), np.zeros(101)))
    dists = []
    for cut in (cut_static, 17.5, cut_dynamic):
        steps = [0.0, np.spacing(cut), -np.spacing(cut), 4 * np.spacing(cut), -4 * np.spacing(cut), 1e-9, -1e-9, 1e-7, -1e-7,
                 5e-5, -5e-5, 3e-4, -3e-4]
        dists += [cut + s for s in steps]
    loc = np.zeros((256, 3))
    vel, wp = w.vel[:256] * 0.0, w.next_waypoint[:256]
    k = 0
    for d in dists:                                       # along x, along y, and on the 3-4-5 direction (sqrt is exact)
        for direction in ((1.0, 0.0), (0.0, -1.0), (0.6, 0.8), (-0.8, 0.6)):
            loc[k, :2] = centre + d * np.array(direction)
            k += 1
    loc[k:, :2] = centre + np.random.default_rng(3).uniform(-60, 60, size=(256 - k, 2))
    ctx = native.Context(0)
    ctx.set_params(native.params_from_config(sfm_config, w.step_length))
    ctx.upload_state(loc, vel, wp, w.radius[:256], w.target_speed[:256], np.ones(256, dtype=np.uint8))
    ctx.set_borders([line], centre[None, :], np.array([17.5]))
    ctx.set_obstacles(native.STATIC_OBSTACLE, [centre], [ring])
    ctx.set_obstacles(native.DYNAMIC_OBSTACLE, [centre], [ring], [np.array([3.0, -1.0])])
    scene = G.scene_for(w, sfm_config)
    _, want_b = O.border_force(loc, w.radius[:256], np.ones(256, dtype=np.uint8), [line], centre[None, :], np.array([17.5]),
                               scene.border, False, return_pairs=True)
    _, want_s = O.obstacle_force(loc, vel, w.radius[:256], [centre], [ring], None, scene.static, False, return_pairs=True)
    _, want_d = O.obstacle_force(loc, vel, w.radius[:256], [centre], [ring], [np.array([3.0, -1.0])], scene.dynamic, False,
                                 return_pairs=True)
    inside = np.linalg.norm(loc[:k, :2] - centre, axis=1) < 17.5
    assert inside.any() and (~inside).any() and 0 < len(want_b) < 256          # both sides of every cutoff are populated
    np.testing.assert_array_equal(ctx.enumerate_pairs(native.BORDER), want_b)
    np.testing.assert_array_equal(ctx.enumerate_pairs(native.STATIC_OBSTACLE), want_s)
    np.testing.assert_array_equal(ctx.enumerate_pairs(native.DYNAMIC_OBSTACLE), want_d)


def test_cell_list_forces_do_not_depend_on_pedestrian_order(sfm_config):
    """A pedestrian's border / obstacle force is summed in (cell, item index) order whatever its neighbours in the table
    are: permuting the pedestrians permutes the forces bit for bit; the pair force (tile partial sums regroup) follows to
    the float32 tolerance."""
    w = synth.make_config(2)
    perm = np.random.default_rng(9).permutation(w.n)
    a = make_context(w, sfm_config)
    b = native.Context(0)
    b.set_params(native.params_from_config(sfm_config, w.step_length))
    b.upload_state(w.loc[perm], w.vel[perm], w.next_waypoint[perm], w.radius[perm], w.target_speed[perm], w.mode[perm])
    b.set_borders(w.borders, w.section_center, w.section_length)
    b.set_obstacles(native.STATIC_OBSTACLE, [c for c, _ in w.static_obstacles], [r for _, r in w.static_obstacles])
    veh = w.vehicles_at(0)
    b.set_obstacles(native.DYNAMIC_OBSTACLE, veh[1], veh[5], veh[3])
    for cls in (native.ACCELERATION, native.BORDER, native.STATIC_OBSTACLE, native.DYNAMIC_OBSTACLE):
        np.testing.assert_array_equal(b.force(cls), a.force(cls)[perm])
    fa, fb = a.force(native.PEDESTRIAN)[perm], b.force(native.PEDESTRIAN)
    assert np.abs(fb - fa).max() <= 1e-5 + 1e-4 * np.abs(fa).max()


@pytest.mark.parametrize('use_radius', [False, True])
def test_appendix_b_edge_cases(sfm_config, use_radius):
    """SURVEY.md appendix B side by side (tests/appendix_b.py -- the oracle is pinned to the imported reference on the same
    scene in tests/test_oracle_vs_reference.py): argmin ties, pedestrians exactly at a cutoff, theta = -pi / +pi, zero
    distance to a border / ring point, masked modes, waypoint reached, zero target speed, |v'| = 0, a pedestrian 3 km away."""
    from tests import appendix_b
    cfg = dict(sfm_config, use_ped_radius=use_radius)
    w = appendix_b.scene()
    ctx = make_context(w, cfg)
    scene = G.scene_for(w, cfg)
    dyn, dyn_vel = G.dyn_for(w, 0)
    with np.errstate(all='ignore'):
        want = _oracle_classes(w, cfg)
        _, risk = O.pedestrian_force(w.loc, w.vel, w.radius, scene.ped, use_radius, return_risk=True)
        _, pairs_b = O.border_force(w.loc, w.radius, w.mode, w.borders, w.section_center, w.section_length, scene.border,
                                    use_radius, return_pairs=True)
        _, pairs_s = O.obstacle_force(w.loc, w.vel, w.radius, [c for c, _ in w.static_obstacles],
                                      [r for _, r in w.static_obstacles], None, scene.static, use_radius, return_pairs=True)
        _, pairs_d = O.obstacle_force(w.loc, w.vel, w.radius, [c for c, _ in dyn], [r for _, r in dyn], dyn_vel,
                                      scene.dynamic, use_radius, return_pairs=True)
        want_loc, want_vel, want_f = O.step(scene, w.loc, w.vel, w.next_waypoint, w.radius, w.target_speed, w.mode, dyn,
                                            dyn_vel)
    np.testing.assert_array_equal(ctx.enumerate_pairs(native.BORDER), pairs_b)
    np.testing.assert_array_equal(ctx.enumerate_pairs(native.STATIC_OBSTACLE), pairs_s)
    np.testing.assert_array_equal(ctx.enumerate_pairs(native.DYNAMIC_OBSTACLE), pairs_d)
    for cls, name in ((native.ACCELERATION, 'acceleration_force'), (native.BORDER, 'border_force'),
                      (native.STATIC_OBSTACLE, 'static_obstacle_force'), (native.DYNAMIC_OBSTACLE, 'dynamic_obstacle_force')):
        np.testing.assert_allclose(ctx.force(cls), want[name], rtol=1e-11, atol=1e-11, err_msg=name)
    got = ctx.force(native.PEDESTRIAN)
    assert_forces_close(got, want['pedestrian_force'], risk=risk, name='appendix-b pairs')
    assert not ctx.force(native.BORDER)[[5, 8, 9]].any() and not ctx.force(native.STATIC_OBSTACLE)[5].any()
    assert not got[2].any()                                        # 3 km from everybody: underflows to exactly zero
    ctx.step(1, integrate_positions=True)
    loc, vel = ctx.download_state()
    tol = 0.05 * (1e-5 + 1e-4 * np.abs(want_f) + risk[:, None])   # dv = dt * dF
    assert (np.abs(vel - want_vel) <= tol + 1e-12).all()
    assert not vel[1].any() and not vel[2].any()                   # target speed 0; |v'| = 0
    np.testing.assert_array_equal(loc[2], w.loc[2])
