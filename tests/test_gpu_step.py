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


def test_cfg2_10_step_trajectory_against_reference_golden(sfm_config):
    """BASELINE.json configs[1] as a run: ten ticks of N = 4,096 with all five forces and moving vehicles against the
    trajectory the imported reference produced.  Stated tolerance: position deviation <= 2e-6 m after one tick, median
    <= 1e-5 m and max <= 1e-3 m after ten (float32 pair forces feeding a dense, chaotic crowd)."""
    w = synth.make_config(2)
    g = G.load('cfg2_trajectory.npz', w)
    ctx = make_context(w, sfm_config)
    steps, rows = list(g['steps']), g['rows']
    for step in range(max(steps)):
        set_vehicles(ctx, w, step)
        ctx.step(1, integrate_positions=True)
        if step + 1 in steps:
            k = steps.index(step + 1)
            loc, vel = ctx.download_state()
            dev = np.linalg.norm(loc[rows] - g['loc'][k], axis=1)
            dv = np.linalg.norm(vel[rows] - g['vel'][k], axis=1)
            limit = {1: 2e-6, 5: 1e-4, 10: 1e-3}[step + 1]
            assert dev.max() <= limit, (step + 1, dev.max())
            if step + 1 == 10:
                assert np.median(dev) <= 1e-5 and np.median(dv) <= 1e-4, (np.median(dev), np.median(dv))


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


def test_two_rank_partition_emulated_on_one_gpu(sfm_config):
    """Row partition + all-gather layout, emulated with two contexts on one GPU (the exchange is a device copy of each
    rank's staged block into the other's gather buffer -- what NCCL's all-gather does across GPUs)."""
    import torch
    from sfm_b200 import engine
    w = synth.make_config(2, n=3000)                                   # ragged: 1500 + 1500 rows in 1536-row blocks
    whole = make_context(w, sfm_config)
    bounds = engine.partition_rows(w.n, 2)
    rows_pad = engine.padded_rows(bounds)
    ranks = []
    for r in range(2):
        lo, hi = int(bounds[r]), int(bounds[r + 1])
        ctx = native.Context(0)
        ctx.set_params(native.params_from_config(sfm_config, w.step_length))
        ctx.set_partition(2, r, rows_pad)
        ctx.upload_state(w.loc[lo:hi], w.vel[lo:hi], w.next_waypoint[lo:hi], w.radius[lo:hi], w.target_speed[lo:hi],
                         w.mode[lo:hi])
        ctx.set_borders(w.borders, w.section_center, w.section_length)
        ctx.set_obstacles(native.STATIC_OBSTACLE, [c for c, _ in w.static_obstacles], [r_ for _, r_ in w.static_obstacles])
        veh = w.vehicles_at(0)
        ctx.set_obstacles(native.DYNAMIC_OBSTACLE, veh[1], veh[5], veh[3])
        ctx.stage()
        ctx.synchronize()
        ptr, per_rank = ctx.gather_buffer()
        view = torch.as_tensor(engine._DeviceView(ptr, per_rank // 4 * 2), device='cuda:0')
        ranks.append((ctx, view, per_rank // 4))

    def exchange():
        for ctx, _, _ in ranks:
            ctx.synchronize()
        for r, (_, view, per) in enumerate(ranks):
            other = ranks[1 - r][1]
            other[r * per:(r + 1) * per].copy_(view[r * per:(r + 1) * per])
        torch.cuda.synchronize()

    def reduce_scatter():
        views = []
        for ctx, _, _ in ranks:
            ctx.synchronize()
            ptr, per = ctx.force_accumulator()
            views.append((torch.as_tensor(engine._DeviceView(ptr, per // 8 * 2, '<i8'), device='cuda:0'), per // 8))
        total = views[0][0] + views[1][0]
        for r, (v, per) in enumerate(views):
            v[r * per:(r + 1) * per].copy_(total[r * per:(r + 1) * per])
        torch.cuda.synchronize()

    exchange()
    for step in range(3):
        whole.step(1, True)
        for ctx, _, _ in ranks:
            ctx.step_begin()
        reduce_scatter()
        for ctx, _, _ in ranks:
            ctx.step_end(True)
        exchange()
    loc_w, vel_w = whole.download_state()
    loc_p = np.concatenate([ranks[r][0].download_state()[0] for r in range(2)])
    vel_p = np.concatenate([ranks[r][0].download_state()[1] for r in range(2)])
    np.testing.assert_allclose(loc_p, loc_w, rtol=0, atol=1e-6)       # split geometry differs -> float32 sum order differs
    np.testing.assert_allclose(vel_p, vel_w, rtol=0, atol=2e-5)


def test_graph_replay_equals_eager_ticks(sfm_config, monkeypatch):
    """With SFM_GRAPH=1 sfm_step replays a captured CUDA graph of the tick after two eager ticks; the trajectory is bit-identical to the
    eager path, and state refreshes / set changes in between are picked up (re-staging, re-capture)."""
    w = synth.make_config(2, n=1536)
    monkeypatch.setenv('SFM_GRAPH', '1')                    # opt-in (not the default: see sfm_api.cu)
    graph = make_context(w, sfm_config)
    monkeypatch.delenv('SFM_GRAPH')
    eager = make_context(w, sfm_config)
    for ctx in (graph, eager):
        ctx.reset_stats()
        ctx.step(12, True)
        loc, vel = ctx.download_state()
        ctx.update_kinematics(loc + 0.01, vel)             # simulator refresh: re-staged outside the graph
        ctx.step(5, True)
        set_vehicles(ctx, w, 7)                            # new vehicle rings: the graph is rebuilt
        ctx.step(5, True)
    a, b = graph.download_state(), eager.download_state()
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1], b[1])
    sg, se = graph.stats(), eager.stats()
    assert se['graph_replays'] == 0 and sg['graph_replays'] >= 12
    assert sg['steps'] == se['steps'] == 22 and sg['launches'] == se['launches']
