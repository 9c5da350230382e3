"""Staged slot order + the pair kernel's run-local path (csrc/k8_order.cuh, k1_sym.cuh), through the C ABI.

Row order belongs to the caller; below the API a rank may stage its rows along a Hilbert curve so that 64-row runs are
compact and far tile pairs are read through the runs' own origins (one subtraction per coordinate and pair).  These tests
turn the reordering on (it is off for a plain context) and hold the results to the same bars as everything else: the
reference's golden vectors, the float64 oracle on evolved states, numpy's zero-safe semantics on degenerate pairs.
"""
import numpy as np
import pytest

from oracle import sfm_oracle as O
from sfm_b200 import native, synth
from tests import golden_util as G
from tests.gpu_util import assert_forces_close, make_context

pytestmark = pytest.mark.gpu


def _local_fraction(stats):
    return stats['local_tile_pairs'] * 65536.0 / max(stats['pair_evaluations'], 1)


@pytest.mark.parametrize('use_radius,z', [(False, 0), (True, 0), (False, 1), (True, 1)])
def test_reordered_staging_against_reference_golden(sfm_config, use_radius, z):
    """cfg2 (N = 4,096), the four kernel instantiations (planar / 3-D x radius off / on) with the rows staged along the
    curve: every row's pair force vs the arrays the imported reference produced -- and the local path really ran."""
    cfg = dict(sfm_config, use_ped_radius=use_radius)
    w = synth.make_config(2, z_spread=0.2 if z else 0.0)
    g = G.load(f'cfg2_forces_r{int(use_radius)}_z{z}.npz', w)
    ctx = make_context(w, cfg)
    ctx.set_reorder_interval(1)
    ctx.reset_stats()
    got = ctx.force(native.PEDESTRIAN)
    stats = ctx.stats()
    order = ctx.slot_order()
    assert sorted(order.tolist()) == list(range(w.n)) and not np.array_equal(order, np.arange(w.n))
    assert stats['local_tile_pairs'] > 0 and _local_fraction(stats) > 0.2, stats
    rows = np.arange(0, w.n, 4)
    _, risk = O.pedestrian_force(w.loc, w.vel, w.radius, G.scene_for(w, cfg).ped, use_radius, rows=rows, return_risk=True)
    assert np.isfinite(got).all()
    assert_forces_close(got[rows], g['F_pedestrian_force'][rows], risk=risk, name='pedestrian_force (reordered)')
    # the other classes never see the staged order
    np.testing.assert_allclose(ctx.force(native.BORDER), g['F_border_force'], rtol=1e-11, atol=1e-11)


def test_slot_order_carries_bitwise_agreement(sfm_config):
    """Two contexts under the same staged order give bit-identical pair forces (integer accumulation above float32 tile
    partials that follow the tile composition); sfm_get_slot_order / sfm_set_slot_order carry the order across."""
    w = synth.make_config(2, n=4096)
    a = make_context(w, sfm_config)
    a.set_reorder_interval(1)
    fa = a.force(native.PEDESTRIAN)
    order = a.slot_order()
    b = make_context(w, sfm_config)                              # reordering off: takes the order it is given
    b.set_slot_order(order)
    fb = b.force(native.PEDESTRIAN)
    np.testing.assert_array_equal(fa, fb)
    assert np.array_equal(b.slot_order(), order)
    c = make_context(w, sfm_config)                              # row order: other tiles, other float32 partial sums --
    fc = c.force(native.PEDESTRIAN)                              # the same forces within the model tolerance, not bitwise
    assert np.array_equal(c.slot_order(), np.arange(w.n))
    assert not np.array_equal(fa, fc)
    pp = O.moussaid_params(sfm_config['pedestrian_force'], O.PED_DEFAULTS)
    rows = np.arange(0, w.n, 8)
    want, risk = O.pedestrian_force(w.loc, w.vel, w.radius, pp, False, rows=rows, return_risk=True)
    assert_forces_close(fa[rows], want, risk=risk, name='staged along the curve')
    assert_forces_close(fc[rows], want, risk=risk, name='staged in row order')
    with pytest.raises(native.SfmError):
        b.set_slot_order(np.zeros(w.n, dtype=np.int32))          # not a permutation


def test_local_path_on_evolved_full_size_state(sfm_config):
    """cfg3 at full size with the order rebuilt every 4 ticks: after 10 integrated ticks the pair force of the float64 state
    on >= 64 rows (incl. the 32 farthest from the origin and the 16 with the closest neighbours) vs the oracle, nearly every
    tile pair on the local path, momentum conserved by the integer accumulators."""
    from scipy.spatial import cKDTree
    w = synth.make_config(3)
    ctx = make_context(w, sfm_config)
    ctx.set_reorder_interval(4)
    ctx.step(10, True)
    loc, vel = ctx.download_state()
    centre = np.round((loc[:, :2].min(axis=0) + loc[:, :2].max(axis=0)) * 0.5)
    far = np.argsort(-np.abs(loc[:, :2] - centre).max(axis=1))[:32]
    nn = cKDTree(loc[:, :2]).query(loc[:, :2], k=2)[0][:, 1]
    rows = np.unique(np.concatenate([np.linspace(0, w.n - 1, 32).astype(np.int64), far, np.argsort(nn)[:16]]))
    pp = O.moussaid_params(sfm_config['pedestrian_force'], O.PED_DEFAULTS)
    want, risk = O.pedestrian_force(loc, vel, w.radius, pp, False, rows=rows, chunk=16, return_risk=True)
    ctx.reset_stats()
    got = ctx.force(native.PEDESTRIAN)
    stats = ctx.stats()
    assert _local_fraction(stats) > 0.85, stats
    assert stats['fixup_rows'] == 0
    assert_forces_close(got[rows], want, risk=risk, name='pedestrian_force, local path, 10 ticks of cfg3')
    total, scale = np.abs(got.sum(axis=0)), np.abs(got).sum(axis=0)
    assert (total[:2] <= 1e-6 * scale[:2]).all()


def test_vanishing_interaction_vector_between_far_tiles(sfm_config):
    """|D| -> 0 (lambda (v_i - v_j) = -e) can happen at any distance, also between tiles the local path reads.  Whether the
    float32 |D| comes out exactly 0 (NaN: both rows go to the repair path, as on the double-single path) or as a few ulp
    (B -> 0, the contribution underflows to 0), the rows must equal numpy's -- which has |D| = 0 exactly: t = 0, B = 0,
    exp(-inf) = 0."""
    w = synth.make_config(2, n=4096)
    loc, vel = w.loc.copy(), w.vel.copy()
    i = int(np.argmin(loc[:, 0] + loc[:, 1]))
    j = int(np.argmax(loc[:, 0] + loc[:, 1]))
    loc[j] = loc[i] + np.array([40.0, 0.0, 0.0])
    vel[i], vel[j] = np.zeros(3), np.array([0.5, 0.0, 0.0])
    w.loc, w.vel = loc, vel
    ctx = make_context(w, sfm_config)
    ctx.set_reorder_interval(1)
    ctx.reset_stats()
    got = ctx.force(native.PEDESTRIAN)
    stats = ctx.stats()
    with np.errstate(all='ignore'):
        want, risk = O.pedestrian_force(loc, vel, w.radius, G.scene_for(w, sfm_config).ped, False, return_risk=True)
    assert np.isfinite(want).all() and np.isfinite(got).all()
    assert stats['local_tile_pairs'] > 0 and stats['fixup_rows'] <= 4, stats
    assert_forces_close(got, want, risk=risk, name='|D| -> 0 across far tiles')


def test_sign_of_zero_angle_on_the_local_path(sfm_config):
    """epsilon = 0 with a standing crowd (theta' = 0 exactly, np.sign(0) = 0, forces.py:108) under the reordered staging."""
    cfg = dict(sfm_config, pedestrian_force=dict(sfm_config['pedestrian_force'], epsilon=0.0))
    w = synth.make_config(2)
    vel = np.zeros_like(w.vel)
    vel[::7] = w.vel[::7]
    ctx = native.Context(0)
    ctx.set_params(native.params_from_config(cfg, w.step_length))
    ctx.set_reorder_interval(1)
    ctx.upload_state(w.loc, vel, w.next_waypoint, w.radius, w.target_speed, w.mode)
    pp = O.moussaid_params(cfg['pedestrian_force'], O.PED_DEFAULTS)
    want, risk = O.pedestrian_force(w.loc, vel, w.radius, pp, False, return_risk=True)
    ctx.reset_stats()
    got = ctx.force(native.PEDESTRIAN)
    stats = ctx.stats()
    assert stats['local_tile_pairs'] > 0 and stats['fixup_rows'] == 0
    assert_forces_close(got, want, risk=risk, name='epsilon = 0, local path')


def test_periodic_reordering_keeps_rows_and_trajectories(sfm_config):
    """12 ticks with the order rebuilt every 3 (all five forces), against the same ticks in row order: rows stay the
    caller's rows, the trajectories agree to float32 rounding of the pair force, and the host-buffer tick and a row-count
    change (fewer rows uploaded) work under a live order."""
    w = synth.make_config(2)
    a, b = make_context(w, sfm_config), make_context(w, sfm_config)
    a.set_reorder_interval(3)
    a.step(12, True)
    b.step(12, True)
    la, va = a.download_state()
    lb, vb = b.download_state()
    np.testing.assert_allclose(la, lb, rtol=0, atol=1e-5)
    np.testing.assert_allclose(va, vb, rtol=0, atol=1e-4)
    assert not np.array_equal(a.slot_order(), np.arange(w.n))
    # host-buffer tick under the live order == the same tick in row order, to rounding
    nv_a, nl_a, nv_b, nl_b = (np.empty((w.n, 3)) for _ in range(4))
    a.tick_host(lb, vb, nv_a, nl_a)
    b.tick_host(lb, vb, nv_b, nl_b)
    np.testing.assert_allclose(nv_a, nv_b, rtol=0, atol=2e-5)
    np.testing.assert_allclose(nl_a, nl_b, rtol=0, atol=1e-6)
    # a smaller crowd uploaded into the same context: the old maps are for another row count and must not be used
    k = 3000
    a.upload_state(w.loc[:k], w.vel[:k], w.next_waypoint[:k], w.radius[:k], w.target_speed[:k], w.mode[:k])
    pp = O.moussaid_params(sfm_config['pedestrian_force'], O.PED_DEFAULTS)
    rows = np.arange(0, k, 8)
    want, risk = O.pedestrian_force(w.loc[:k], w.vel[:k], w.radius[:k], pp, False, rows=rows, return_risk=True)
    got = a.force(native.PEDESTRIAN)
    assert got.shape == (k, 3) and sorted(a.slot_order().tolist()) == list(range(k))
    assert_forces_close(got[rows], want, risk=risk, name='after a row-count change')
