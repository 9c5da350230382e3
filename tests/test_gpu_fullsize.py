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


# ---- parity on EVOLVED states: after integrated ticks positions leave the float32 lattice ------------------------------
def _evolved_rows(loc, n_uniform, n_far):
    """Uniformly spaced rows plus the rows farthest from the staging origin (where float32 staging would be coarsest)."""
    centre = np.round((loc[:, :2].min(axis=0) + loc[:, :2].max(axis=0)) * 0.5)
    far = np.argsort(-np.abs(loc[:, :2] - centre).max(axis=1))[:n_far]
    return np.unique(np.concatenate([np.linspace(0, len(loc) - 1, n_uniform).astype(np.int64), far]))


@pytest.mark.parametrize('cfg_id,ticks,n_uniform,n_far,chunk', [(3, 10, 48, 32, 16), (5, 10, 32, 32, 4)])
def test_pair_force_on_evolved_state(sfm_config, cfg_id, ticks, n_uniform, n_far, chunk):
    """10 integrated ticks, then the pair force of the float64 state the engine itself produced (forces.py:74-117 on
    pedestrian_state.py:17 float64 fields) on >= 64 rows incl. the 32 farthest from the origin: 1e-4 rel + 1e-5 abs."""
    w = synth.make_config(cfg_id)
    ctx = make_context(w, sfm_config) if cfg_id == 3 else None
    if ctx is None:
        ctx = native.Context(0)
        ctx.set_params(native.params_from_config(sfm_config, w.step_length))
        ctx.upload_state(w.loc, w.vel, w.next_waypoint, w.radius, w.target_speed, w.mode)
    ctx.step(ticks, True)
    loc, vel = ctx.download_state()
    assert np.abs(loc - w.loc).max() > 0.1                                   # the crowd really moved ...
    assert (loc[:, :2].astype(np.float32).astype(np.float64) != loc[:, :2]).mean() > 0.9    # ... off the float32 lattice
    rows = _evolved_rows(loc, n_uniform, n_far)
    assert len(rows) >= 60
    pp = O.moussaid_params(sfm_config['pedestrian_force'], O.PED_DEFAULTS)
    want, risk = O.pedestrian_force(loc, vel, w.radius, pp, False, rows=rows, chunk=chunk, return_risk=True)
    got = ctx.force(native.PEDESTRIAN)
    assert_forces_close(got[rows], want, risk=risk, name=f'pedestrian_force after {ticks} ticks of cfg{cfg_id}')
    # the integer accumulators conserve momentum on any state
    total, scale = np.abs(got.sum(axis=0)), np.abs(got).sum(axis=0)
    assert (total[:2] <= 1e-6 * scale[:2]).all()


@pytest.mark.parametrize('offset', [(0.0, 0.0), (512.123456789, -498.87654321), (98765.4321, 123456.789)])
@pytest.mark.parametrize('explicit_origin', [False, True])
def test_pair_force_independent_of_origin(sfm_config, offset, explicit_origin):
    """Non-float32-exact coordinates far from (0, 0), with and without sfm_set_origin: the staged (hi, lo) pairs make the
    pair force independent of where the crowd sits (every row of N = 4,096 against the oracle)."""
    w = synth.make_config(2)
    rng = np.random.default_rng(7)
    loc = w.loc.copy()
    loc[:, :2] += rng.uniform(-0.03, 0.03, size=(w.n, 2)) + np.asarray(offset)
    vel = w.vel + rng.normal(0, 1e-3, size=w.vel.shape) * [1, 1, 0]
    ctx = native.Context(0)
    ctx.set_params(native.params_from_config(sfm_config, w.step_length))
    if explicit_origin:
        ctx.set_origin(float(np.round(offset[0])), float(np.round(offset[1])), 0.0)
    ctx.upload_state(loc, vel, w.next_waypoint, w.radius, w.target_speed, w.mode)
    pp = O.moussaid_params(sfm_config['pedestrian_force'], O.PED_DEFAULTS)
    want, risk = O.pedestrian_force(loc, vel, w.radius, pp, False, return_risk=True)
    assert_forces_close(ctx.force(native.PEDESTRIAN), want, risk=risk, name=f'offset {offset}')


def test_sign_of_zero_angle(sfm_config):
    """theta' == 0 gives np.sign(0) = 0, i.e. no tangential force (forces.py:108).  With epsilon = 0
    every pair of a standing crowd has theta' = 0 exactly; the fast path must not hand out +-f_theta there."""
    cfg = dict(sfm_config, pedestrian_force=dict(sfm_config['pedestrian_force'], epsilon=0.0))
    w = synth.make_config(2)
    vel = np.zeros_like(w.vel)
    vel[::7] = w.vel[::7]                       # a few walkers among a standing crowd
    ctx = native.Context(0)
    ctx.set_params(native.params_from_config(cfg, w.step_length))
    ctx.upload_state(w.loc, vel, w.next_waypoint, w.radius, w.target_speed, w.mode)
    pp = O.moussaid_params(cfg['pedestrian_force'], O.PED_DEFAULTS)
    # `risk` counts 2 |f_theta| of pairs near the discontinuity: exactly 0 for the standing pairs (their f_theta is 0), so
    # they get no slack -- a fast path that ignored sign(0) = 0 would miss by A exp(-d/B) per pair
    want, risk = O.pedestrian_force(w.loc, vel, w.radius, pp, False, return_risk=True)
    got = ctx.force(native.PEDESTRIAN)
    assert_forces_close(got, want, risk=risk, name='epsilon = 0')
    standing = np.ones(w.n, dtype=bool)
    standing[::7] = False
    assert np.abs(want[standing]).max() > 1e-2 and ctx.stats()['fixup_rows'] == 0


def test_pair_force_translation_invariance_is_exact(sfm_config):
    """The staged hi parts live on a 2^-6 m lattice and the kernel only ever forms differences: moving the whole crowd by a
    lattice vector (here 4,096 m and -8,192 m, far beyond where a single float32 could hold 1e-5 m) leaves every bit of the
    pair force unchanged -- and so does moving the staging origin."""
    w = synth.make_config(2)
    rng = np.random.default_rng(5)
    loc = w.loc.copy()
    loc[:, :2] += rng.uniform(-0.03, 0.03, size=(w.n, 2))              # off the float32 lattice ...
    loc[:, :2] = np.rint(loc[:, :2] * 2.0 ** 30) / 2.0 ** 30            # ... but on 2^-30 m, so that float64 can shift it exactly
    assert (loc[:, :2].astype(np.float32).astype(np.float64) != loc[:, :2]).mean() > 0.9
    shift = np.array([4096.0, -8192.0, 0.0])
    assert np.array_equal((loc + shift) - shift, loc)                   # the shift itself is exact in float64
    out = []
    for offset, origin in ((0.0 * shift, None), (shift, None), (shift, (4000.0, -8000.0, 0.0)), (0.0 * shift, (-77.0, 13.0, 0.0))):
        ctx = native.Context(0)
        ctx.set_params(native.params_from_config(sfm_config, w.step_length))
        if origin is not None:
            ctx.set_origin(*origin)
        ctx.upload_state(loc + offset, w.vel, w.next_waypoint + offset, w.radius, w.target_speed, w.mode)
        out.append(ctx.force(native.PEDESTRIAN))
        ctx.close()
    for other in out[1:]:
        np.testing.assert_array_equal(other, out[0])
