"""Host-side mirror of the reference call surface: stateutils, PedState, PedModeManager, check_traffic (CPU only).

Where /root/reference is present every function is compared with the reference's own on random inputs; the
structural expectations (shapes, zero handling, view aliasing) are asserted everywhere.
"""
import numpy as np
import pytest

import ped_mode_manager
import pedestrian_state
import stateutils
from oracle import ref_loader


@pytest.fixture(scope='module')
def ref():
    return ref_loader.load() if ref_loader.available() else None


def test_normalize_zero_safe():
    v, n = stateutils.normalize(np.array([[3.0, 4.0], [0.0, 0.0]]))
    np.testing.assert_array_equal(v, [[0.6, 0.8], [0.0, 0.0]])
    np.testing.assert_array_equal(n, [5.0, 0.0])


def test_all_diffs_orientation_and_shape():
    a = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 2.0]])
    d = stateutils.all_diffs(a)
    assert d.shape == (3, 2, 2)
    np.testing.assert_array_equal(d[0], [[1.0, 0.0], [0.0, 2.0]])      # a[j] - a[0], j = 1, 2
    np.testing.assert_array_equal(d[1], [[-1.0, 0.0], [-1.0, 2.0]])    # j = 0, 2 (i skipped)
    s = stateutils.all_sums(np.array([1.0, 2.0, 4.0]))
    np.testing.assert_array_equal(s, [[3.0, 5.0], [3.0, 6.0], [5.0, 6.0]])


def test_stateutils_match_reference(ref):
    if ref is None:
        pytest.skip('reference tree not present')
    rng = np.random.default_rng(7)
    a3, b3 = rng.normal(size=(17, 3)), rng.normal(size=(17, 3))
    a3[4] = 0.0
    np.testing.assert_array_equal(stateutils.all_diffs(a3), ref.stateutils.all_diffs(a3))
    np.testing.assert_array_equal(stateutils.all_sums(a3), ref.stateutils.all_sums(a3))
    # 1-D input: the reference's all_diffs raises IndexError (stateutils.py:44), all_sums works (forces.py:81 uses it)
    np.testing.assert_array_equal(stateutils.all_sums(a3[:, 0]), ref.stateutils.all_sums(a3[:, 0]))
    np.testing.assert_array_equal(stateutils.all_diffs(a3, keep_dims=False), ref.stateutils.all_diffs(a3, keep_dims=False))
    np.testing.assert_array_equal(stateutils.all_diffs(a3, remove_diagonal=False),
                                  ref.stateutils.all_diffs(a3, remove_diagonal=False))
    for m, t in zip(stateutils.normalize(a3), ref.stateutils.normalize(a3)):
        np.testing.assert_array_equal(m, t)
    np.testing.assert_array_equal(stateutils.angle_diff_2d(a3, b3), ref.stateutils.angle_diff_2d(a3, b3))
    assert stateutils.angle_diff_2d(a3[0], b3[0]) == ref.stateutils.angle_diff_2d(a3[0], b3[0])
    np.testing.assert_array_equal(stateutils.cap_velocity(a3, np.full(17, 0.7)), ref.stateutils.cap_velocity(a3, np.full(17, 0.7)))
    state = np.zeros(17, dtype=[('loc', 'f8', (3,)), ('vel', 'f8', (3,)), ('next_waypoint', 'f8', (3,))])
    state['loc'], state['vel'], state['next_waypoint'] = a3, b3, rng.normal(size=(17, 3))
    state['next_waypoint'][2] = state['loc'][2]
    np.testing.assert_array_equal(stateutils.desired_directions(state), ref.stateutils.desired_directions(state))
    np.testing.assert_array_equal(stateutils.speeds(state), ref.stateutils.speeds(state))


def _walk(manager_cls, mode_cls):
    m = manager_cls('p0', 1.2, mode_cls.WALKING_SIDEWALK, 1.5, 1.0)
    log = []
    for t, action in enumerate([None, mode_cls.CROSSING_ROAD, mode_cls.CROSSING_ROAD, mode_cls.WALKING_SIDEWALK,
                                mode_cls.WALKING_SIDEWALK, mode_cls.IDLE, None, None, None, None, None, None]):
        if action is not None:
            m.set_mode(action)
        m.tick(float(t))
        log.append((int(m.current_mode), float(m.target_speed)))
    return log


def test_mode_machine_sequence(ref):
    mine = _walk(ped_mode_manager.PedModeManager, ped_mode_manager.PedMode)
    assert mine[1] == (4, 0.0)                      # sidewalk -> crossing goes through CHECKING_TRAFFIC at speed 0
    assert mine[2][0] == 2 and abs(mine[2][1] - 1.8) < 1e-12
    assert mine[3][0] == 3                          # crossing -> sidewalk goes through ROAD_TO_SIDEWALK
    assert mine[5] == (0, 0.0) and mine[-1][0] == 1  # idle for waiting_time, then walking again
    if ref is not None:
        assert mine == _walk(ref.ped_mode_manager.PedModeManager, ref.ped_mode_manager.PedMode)


def test_pedstate_surface(ref):
    cfg = {}
    mode = ped_mode_manager.PedMode
    mk = lambda name, speed: ped_mode_manager.PedModeManager(name, speed, mode.WALKING_SIDEWALK, 1.5, 1.0)   # noqa: E731
    peds = pedestrian_state.PedState(cfg)
    for k in range(4):
        peds.add_pedestrian((f'p{k}', 10 + k, [k, 0.0, 1.0], [0.1, 0.2, 0.0], [5.0, 5.0, 1.0], mk(f'p{k}', 1.0 + k),
                             0.3, 1.0 + k))
    assert peds.size() == 4 and peds.state.dtype.names == ('name', 'id', 'loc', 'vel', 'next_waypoint', 'mode',
                                                           'radius', 'target_speed')
    peds.update_state(12, [9.0, 9.0, 1.0], [1.0, 0.0, 0.0])
    np.testing.assert_array_equal(peds.loc()[2], [9.0, 9.0, 1.0])
    peds.update_states(np.array([13, 10]), np.array([[1.0, 1, 1], [2.0, 2, 2]]), np.zeros((2, 3)))
    np.testing.assert_array_equal(peds.loc()[[3, 0]], [[1.0, 1, 1], [2.0, 2, 2]])
    peds.update_next_waypoint('p1', ([7.0, 7.0, 1.0], True))
    assert peds.mode()[1].current_mode == mode.CHECKING_TRAFFIC
    peds.apply_current_mode()
    np.testing.assert_array_equal(peds.target_speed(), [1.0, 0.0, 3.0, 4.0])
    np.testing.assert_array_equal(peds.max_speed(), np.array([1.0, 0.0, 3.0, 4.0]) * 1.3)
    np.testing.assert_array_equal(peds.mode_codes(), [1, 4, 1, 1])
    peds.record_current_state(0.5)
    assert peds.get_all_states()[0.5]['mode'][1] == mode.CHECKING_TRAFFIC
    view = peds.state[['id', 'vel']]
    view['vel'] = np.ones((4, 3))
    assert peds.vel().sum() == 12.0                  # multi-field index is a view: writes reach state['vel']
    peds.remove_pedestrian('p0')
    assert list(peds.name()) == ['p1', 'p2', 'p3']
    cols = peds.device_columns()
    assert [c.dtype for c in cols] == [np.float64] * 5 + [np.uint8] and all(c.flags.c_contiguous for c in cols)
    assert pedestrian_state.PedState({'max_speed_factor': 2.0, 'max_speed_multiplier': 9.0}).max_speed_factor == 2.0


def test_check_traffic_gap_acceptance():
    from check_traffic import check_traffic
    mode = ped_mode_manager.PedModeManager('p', 1.0, ped_mode_manager.PedMode.CHECKING_TRAFFIC, 1.0, 1.0)
    ped = {'loc': np.array([0.0, -5.0, 0.0]), 'next_waypoint': np.array([0.0, 5.0, 0.0]), 'mode': mode}
    ring = np.zeros((4, 2))
    extents = np.array([[2.0, 1.0]])
    near = [(np.array([-10.0, 0.0]), ring)]
    assert check_traffic(ped, near, np.array([[2.0, 0.0]]), extents) is False          # arrives with the pedestrian
    assert check_traffic(ped, [(np.array([-200.0, 0.0]), ring)], np.array([[2.0, 0.0]]), extents) is True   # too far
    assert check_traffic(ped, near, np.array([[0.0, 0.0]]), extents) is True           # parked
    mode.crossing_safety_margin = -1.0
    assert check_traffic(ped, near, np.array([[2.0, 0.0]]), extents) is True           # negative margin: no check


def test_mode_table_equals_standalone_machines():
    """A PedModeManager adopted into a ModeTable behaves exactly like a standalone one: same modes, speeds and wake-up
    times under random requests and ticks, through the objects and through the table's vectorised tick."""
    rng = np.random.default_rng(11)
    n = 200
    PM = ped_mode_manager.PedMode

    def make():
        return [ped_mode_manager.PedModeManager(f'p{i}', 1.0 + 0.01 * i, PM(int(rng2.integers(0, 5))), 1.5, 1.0)
                for i in range(n)]
    rng2 = np.random.default_rng(5)
    alone = make()
    rng2 = np.random.default_rng(5)
    tabled = make()
    table = ped_mode_manager.ModeTable(tabled)
    assert all(m._table is table for m in tabled)
    for step in range(300):
        t = 0.05 * step
        for k in rng.integers(0, n, size=6):
            wanted = PM(int(rng.integers(0, 5)))
            alone[k].set_mode(wanted)
            tabled[k].set_mode(wanted)
        for m in alone:
            m.tick(t)
        table.tick(t)
        got = [(int(m.current_mode), float(m.target_speed), float(m.next_mode_time), float(m.sim_time)) for m in tabled]
        want = [(int(m.current_mode), float(m.target_speed), float(m.next_mode_time), float(m.sim_time)) for m in alone]
        assert got == want, step
    assert table.version > 0
    table.release()                                 # the objects take their values back and keep working
    assert all(m._table is None for m in tabled)
    assert [int(m.current_mode) for m in tabled] == [int(m.current_mode) for m in alone]
    tabled[0].set_mode(PM.IDLE)
    assert tabled[0].current_mode == PM.IDLE and tabled[0].target_speed == 0


def test_ped_state_adopts_and_tracks_mode_objects():
    cfg = {}
    ps = pedestrian_state.PedState(cfg)
    PM = ped_mode_manager.PedMode
    modes = [ped_mode_manager.PedModeManager(f'p{i}', 1.2, PM.WALKING_SIDEWALK, 1.5, 1.0) for i in range(50)]
    ps.add_pedestrians([f'p{i}' for i in range(50)], np.arange(50), np.zeros((50, 3)), np.zeros((50, 3)),
                       np.zeros((50, 3)), modes, np.full(50, 0.3), np.full(50, 1.2))
    table = ps.mode_table()
    assert table is not None and ps.mode_table() is table           # same objects -> same table
    modes[7].set_mode(PM.CROSSING_ROAD)
    assert ps.mode_codes()[7] == PM.CHECKING_TRAFFIC
    ps.apply_current_mode()
    assert ps.state['target_speed'][7] == 0.0 and ps.state['target_speed'][8] == 1.2
    ps.remove_pedestrian('p7')                                       # the row set changed: a new table, p7 stands alone
    table2 = ps.mode_table()
    assert table2 is not table and modes[7]._table is None and modes[7].current_mode == PM.CHECKING_TRAFFIC
    assert modes[8]._table is table2 and modes[8]._row == 7
    ps.state['mode'][3] = PM.WALKING_SIDEWALK                        # a plain enum in the column: no table, generic path
    assert ps.mode_table() is None and modes[4]._table is None
    assert ps.mode_codes()[3] == PM.WALKING_SIDEWALK
