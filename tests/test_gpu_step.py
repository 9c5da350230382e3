"""GPU parity of the fused tick (forces + clamp + Euler) against the reference's cfg1 golden trajectory."""
import numpy as np
import pytest

from oracle import sfm_oracle as O
from sfm_b200 import native, synth
from tests import golden_util as G
from tests.gpu_util import make_context, set_vehicles

pytestmark = pytest.mark.gpu


def test_cfg1_100_step_trajectory(sfm_config):
    """BASELINE.json configs[0]: 100 ticks, all five forces.  Stated tolerance (float32 pair forces feeding a chaotic
    system): position deviation median <= 1e-4 m and max <= 1e-2 m after 100 steps,
    and <= 1e-5 m after the first step."""
    w = synth.make_config(1)
    g = G.load('cfg1_trajectory.npz', w)
    ctx = make_context(w, sfm_config)
    for step in range(100):
        set_vehicles(ctx, w, step)
        ctx.step(1, integrate_positions=True)
        if step in (0, 9, 49, 99):
            loc, vel = ctx.download_state()
            dev = np.linalg.norm(loc - g['loc'][step + 1], axis=1)
            limit_max = {0: 1e-5, 9: 1e-4, 49: 2e-3, 99: 1e-2}[step]
            assert dev.max() <= limit_max, (step, dev.max())
            if step == 99:
                assert np.median(dev) <= 1e-4
            speed = np.linalg.norm(vel, axis=1)
            assert (speed <= w.target_speed * 1.3 * (1 + 1e-12)).all()


def test_one_step_matches_oracle_velocity_update(sfm_config):
    """With the pedestrian force off every remaining kernel is float64 in numpy's operation order: the whole step
    (forces, clamp, Euler) must then agree with the oracle to rounding."""
    enable = dict(sfm_config['forces'], pedestrian_force=False)
    cfg = dict(sfm_config, forces=enable)
    w = synth.make_config(2, n=1024)
    ctx = make_context(w, cfg)
    ctx.step(1, integrate_positions=True)
    loc, vel = ctx.download_state()
    dyn, dyn_vel = G.dyn_for(w, 0)
    want_loc, want_vel, want_f = O.step(G.scene_for(w, cfg), w.loc, w.vel, w.next_waypoint, w.radius, w.target_speed,
                                        w.mode, dyn, dyn_vel)
    np.testing.assert_allclose(ctx.download_force(), want_f, rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(vel, want_vel, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(loc, want_loc, rtol=1e-13, atol=1e-13)


def test_velocity_only_tick_and_host_tick(sfm_config):
    """CARLA-coupled mode: positions are not integrated (pedestrian_simulation.py:117-124); host-buffer tick agrees."""
    w = synth.make_config(2, n=2048)
    a = make_context(w, sfm_config)
    a.step(1, integrate_positions=False)
    loc_a, vel_a = a.download_state()
    np.testing.assert_array_equal(loc_a, w.loc)
    b = make_context(w, sfm_config)
    new_vel = np.empty((w.n, 3))
    b.tick_host(np.ascontiguousarray(w.loc), np.ascontiguousarray(w.vel), new_vel)
    np.testing.assert_array_equal(new_vel, vel_a)


def test_zero_target_speed_stops(sfm_config):
    w = synth.make_config(1)
    w.target_speed = np.zeros(w.n)
    ctx = make_context(w, sfm_config)
    ctx.step(1)
    _, vel = ctx.download_state()
    assert not vel.any()                                      # stateutils.py:20-23 with s = 0
