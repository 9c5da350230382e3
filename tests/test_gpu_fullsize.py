"""-m gpu: parity at BASELINE.json's full sizes, where the float64 oracle can only afford a row sample -- row-sample
comparisons plus size-independent properties of the model (Newton's third law for the pair force, exact agreement of the
neighbour enumeration on the sampled pedestrians)."""
import numpy as np
import pytest

from oracle import sfm_oracle as O
from sfm_b200 import native, synth
from tests.gpu_util import assert_forces_close, make_context

pytestmark = pytest.mark.gpu


def test_cfg3_full_size_row_sample_and_third_law(sfm_config):
    """cfg3: N = 65,536 + 1.05 M border points + 50 k obstacle points.  128 sampled rows of every force class and of the
    new velocities against the oracle; the enumeration of the sampled pedestrians bit for bit; sum of all pair forces ~ 0."""
    w = synth.make_config(3)
    ctx = make_context(w, sfm_config)
    rows = np.linspace(0, w.n - 1, 128).astype(np.int64)
    scene = O.Scene(sfm_config, w.step_length, w.borders, w.section_center, w.section_length, w.static_obstacles)
    want = O.forces_by_class(scene, w.loc, w.vel, w.next_waypoint, w.radius, w.target_speed, w.mode, rows=rows, chunk=16)
    _, risk = O.pedestrian_force(w.loc, w.vel, w.radius, scene.ped, scene.use_ped_radius, rows=rows, chunk=16,
                                 return_risk=True)
    f_ped = ctx.force(native.PEDESTRIAN)
    assert_forces_close(f_ped[rows], want['pedestrian_force'], risk=risk, name='pedestrian_force')
    np.testing.assert_allclose(ctx.force(native.ACCELERATION)[rows], want['acceleration_force'], rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(ctx.force(native.BORDER)[rows], want['border_force'], rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(ctx.force(native.STATIC_OBSTACLE)[rows], want['static_obstacle_force'], rtol=1e-10, atol=1e-10)
    # Newton's third law: f_ji = -f_ij, so the pair forces of the whole crowd cancel (exactly, in the integer accumulators,
    # for every tile pair off the diagonal; to float32 rounding inside the 256 diagonal tiles)
    total, scale = np.abs(f_ped.sum(axis=0)), np.abs(f_ped).sum(axis=0)
    assert (total[:2] <= 1e-6 * scale[:2]).all() and total[2] == 0.0, (total, scale)
    # neighbour enumeration of the sampled pedestrians (full sets: 5,000 sections, 2,500 obstacles), bit for bit
    sel = np.zeros(w.n, dtype=bool)
    sel[rows] = True
    remap = np.full(w.n, -1)
    remap[rows] = np.arange(len(rows))
    for cls, ref in ((native.BORDER, lambda: O.border_force(w.loc[rows], w.radius[rows], w.mode[rows], w.borders,
                                                             w.section_center, w.section_length, scene.border, False,
                                                             return_pairs=True)[1]),
                     (native.STATIC_OBSTACLE, lambda: O.obstacle_force(w.loc[rows], w.vel[rows], w.radius[rows],
                                                                       [c for c, _ in w.static_obstacles],
                                                                       [r for _, r in w.static_obstacles], None,
                                                                       scene.static, False, return_pairs=True)[1])):
        got = ctx.enumerate_pairs(cls, capacity=1 << 23)
        got = got[sel[got[:, 0]]]
        got[:, 0] = remap[got[:, 0]]
        got = got[np.lexsort((got[:, 1], got[:, 0]))]
        assert np.array_equal(got, ref()), cls
    # one full tick for the sampled rows
    ctx.step(1, True)
    _, vel = ctx.download_state()
    F = O.total_force(want, len(rows))
    v_want = O.new_velocities(w.vel[rows], F, w.target_speed[rows], scene.dt, scene.max_speed_factor)
    assert np.abs(vel[rows] - v_want).max() < 2e-5


def test_cfg5_full_size_row_sample_and_third_law(sfm_config):
    """cfg5: N = 1,048,576.  16 sampled rows of the pair force against the oracle (1e-4 rel + 1e-5 abs) and the third law
    over the whole crowd.  (8-GPU vs 1-GPU bitwise trajectory agreement: profiles/cfg5_agreement_r1.json.)"""
    w = synth.make_config(5)
    ctx = native.Context(0)
    ctx.set_params(native.params_from_config(sfm_config, w.step_length))
    ctx.upload_state(w.loc, w.vel, w.next_waypoint, w.radius, w.target_speed, w.mode)
    f_ped = ctx.force(native.PEDESTRIAN)
    rows = np.linspace(0, w.n - 1, 16).astype(np.int64)
    pp = O.moussaid_params(sfm_config['pedestrian_force'], O.PED_DEFAULTS)
    want, risk = O.pedestrian_force(w.loc, w.vel, w.radius, pp, False, rows=rows, chunk=4, return_risk=True)
    assert_forces_close(f_ped[rows], want, risk=risk, name='pedestrian_force at N = 1,048,576')
    total, scale = np.abs(f_ped.sum(axis=0)), np.abs(f_ped).sum(axis=0)
    assert (total[:2] <= 1e-6 * scale[:2]).all() and total[2] == 0.0
    assert ctx.stats()['fixup_rows'] == 0
