"""Oracle vs the committed golden vectors (outputs of the imported reference, see oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import sfm_oracle as O
from sfm_b200 import synth
from tests import golden_util as G


def test_cfg1_trajectory(sfm_config):
    w = synth.make_config(1)
    g = G.load('cfg1_trajectory.npz', w)
    scene = G.scene_for(w, sfm_config)
    loc, vel = w.loc.copy(), w.vel.copy()
    keep = list(g['force_steps'])
    for step in range(100):
        dyn, dyn_vel = G.dyn_for(w, step)
        if step in keep:
            per_class = O.forces_by_class(scene, loc, vel, w.next_waypoint, w.radius, w.target_speed, w.mode, dyn, dyn_vel)
            for name, f in per_class.items():
                np.testing.assert_allclose(f, g[f'F_{name}'][keep.index(step)], rtol=1e-9, atol=1e-10, err_msg=name)
        loc, vel, _ = O.step(scene, loc, vel, w.next_waypoint, w.radius, w.target_speed, w.mode, dyn, dyn_vel)
        np.testing.assert_allclose(loc, g['loc'][step + 1], rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(vel, g['vel'][step + 1], rtol=1e-9, atol=1e-10)


@pytest.mark.parametrize('use_radius,z', [(False, 0), (True, 0), (False, 1), (True, 1)])
def test_cfg2_force_classes(sfm_config, use_radius, z):
    w = synth.make_config(2, z_spread=0.2 if z else 0.0)
    g = G.load(f'cfg2_forces_r{int(use_radius)}_z{z}.npz', w)
    cfg = dict(sfm_config, use_ped_radius=use_radius)
    scene = G.scene_for(w, cfg)
    rows = np.arange(0, w.n, 8) if (use_radius or z) else None           # one variant in full, the others sampled
    dyn, dyn_vel = G.dyn_for(w, 0)
    per_class = O.forces_by_class(scene, w.loc, w.vel, w.next_waypoint, w.radius, w.target_speed, w.mode, dyn, dyn_vel,
                                  rows=rows)
    sel = slice(None) if rows is None else rows
    for name, f in per_class.items():
        np.testing.assert_allclose(f, g[f'F_{name}'][sel], rtol=1e-12, atol=1e-12, err_msg=name)


def test_cfg2_trajectory_first_ticks(sfm_config):
    """The reference's own 10-tick loop at N = 4,096 (tests/golden/cfg2_trajectory.npz): the oracle reproduces the state
    after ticks 1 and 5 (the GPU test covers all ten)."""
    w = synth.make_config(2)
    g = G.load('cfg2_trajectory.npz', w)
    scene = G.scene_for(w, sfm_config)
    loc, vel = w.loc.copy(), w.vel.copy()
    steps, rows = list(g['steps']), g['rows']
    for step in range(5):
        dyn, dyn_vel = G.dyn_for(w, step)
        loc, vel, _ = O.step(scene, loc, vel, w.next_waypoint, w.radius, w.target_speed, w.mode, dyn, dyn_vel)
        if step + 1 in steps:
            k = steps.index(step + 1)
            np.testing.assert_allclose(loc[rows], g['loc'][k], rtol=1e-9, atol=1e-9)
            np.testing.assert_allclose(vel[rows], g['vel'][k], rtol=1e-9, atol=1e-9)
