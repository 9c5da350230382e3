"""Helpers for the -m gpu parity tests: build a device context from a synthetic workload (through the C ABI)."""
import numpy as np

from sfm_b200 import native


def make_context(w, cfg, step=0, device=0, enable=None):
    ctx = native.Context(device)
    ctx.set_params(native.params_from_config(cfg, w.step_length, enable=enable))
    ctx.upload_state(w.loc, w.vel, w.next_waypoint, w.radius, w.target_speed, w.mode)
    if len(w.borders):
        ctx.set_borders(w.borders, w.section_center, w.section_length)
    if len(w.static_obstacles):
        ctx.set_obstacles(native.STATIC_OBSTACLE, [c for c, _ in w.static_obstacles],
                          [r for _, r in w.static_obstacles])
    set_vehicles(ctx, w, step)
    return ctx


def set_vehicles(ctx, w, step):
    veh = w.vehicles_at(step)
    if veh is not None:
        ctx.set_obstacles(native.DYNAMIC_OBSTACLE, veh[1], veh[5], veh[3])


def assert_forces_close(got, want, rtol=1e-4, atol=1e-5, risk=None, name=''):
    """|got - want| <= atol + rtol * |want| per component; rows carrying ill-conditioned pairs (oracle ``risk``: force
    magnitude sitting within 1e-5 rad of the sign / wrap discontinuities) get that magnitude as extra slack."""
    err = np.abs(got - want)
    tol = atol + rtol * np.abs(want)
    if risk is not None:
        tol = tol + risk[:, None]
    bad = err > tol
    if bad.any():
        r, c = np.unravel_index(np.argmax(err - tol), err.shape)
        raise AssertionError(f'{name}: {bad.sum()} components out of tolerance; worst row {r} comp {c}: '
                             f'got {got[r, c]!r} want {want[r, c]!r} err {err[r, c]:.3e} tol {tol[r, c]:.3e}')
