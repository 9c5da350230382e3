"""A hand-built scene that puts the edge cases of SURVEY.md appendix B next to each other (test input, no device code).

Rows (N = 24; rows 16.. are ordinary pedestrians so that every row also carries regular pair forces):
  0   stands exactly on its waypoint              -> desired direction 0, F_acc = -v / tau      (stateutils.py:9-10, :88-90)
  1   target_speed = 0                            -> new velocity exactly 0                      (stateutils.py:20-23)
  2   v = 0, on its waypoint, 3 km from anybody   -> every force underflows to 0, |v'| = 0       (stateutils.py:21)
  3   equidistant from two points of section 0    -> np.argmin takes the first index             (forces.py:154)
  4   equidistant from four points of obstacle 0  -> first index                                  (forces.py:228)
  5   beyond every section length / threshold     -> border and obstacle forces exactly [0, 0]   (forces.py:166-167, :274-275)
  6,7 D antiparallel to d                         -> theta = -pi and +pi, no wrap applied         (stateutils.py:111-112)
  8   CROSSING_ROAD next to section 0             -> border row zeroed                            (forces.py:176-177)
  9   ROAD_TO_SIDEWALK next to section 0          -> border row zeroed
  10  exactly `section_length` from the centre of section 1 (3-4-5 triangle) -> strict '<' excludes it (forces.py:149-150)
  11  exactly `perception_threshold` from the centre of obstacle 1           -> excluded           (forces.py:222-223)
  12  sits on a border point of section 0 (dist = 0) -> normalize gives direction 0               (forces.py:158)
  13  sits on a ring point of obstacle 0             -> d = 0: e = 0, D = lambda (v - u)          (forces.py:231-246)
"""
import numpy as np

from sfm_b200 import synth

ROAD_TO_SIDEWALK = 3                                   # ped_mode_manager.py:4-9


def scene():
    rng = np.random.default_rng(77)
    n = 24
    loc = np.zeros((n, 3))
    vel = np.zeros((n, 3))
    wp = np.zeros((n, 3))
    loc[:, :2] = rng.uniform(20.0, 30.0, size=(n, 2))
    vel[:, :2] = rng.normal(0.0, 1.0, size=(n, 2))
    wp[:, :2] = rng.uniform(0.0, 40.0, size=(n, 2))
    radius = rng.uniform(0.2, 0.4, size=n)
    speed = rng.uniform(1.0, 1.6, size=n)
    mode = np.full(n, synth.WALKING_SIDEWALK, dtype=np.uint8)

    # section 0: points (10 + k, 10), k = 0..20; section 1: short, cutoff 5; section 2: far away
    s0 = np.column_stack((10.0 + np.arange(21.0), np.full(21, 10.0)))
    s1 = np.column_stack((np.full(5, 60.0), 60.0 + 0.5 * np.arange(5.0)))
    s2 = np.column_stack((500.0 + np.arange(10.0), np.full(10, 500.0)))
    borders = [s0, s1, s2]
    centres = np.array([s0[10], s1[2], s2[5]])
    lengths = np.array([25.0, 5.0, 10.0])
    # obstacle 0: square ring of four points around (30, 30); obstacle 1: hexagon at (80, 20)
    ring0 = np.array([[29.0, 29.0], [31.0, 29.0], [31.0, 31.0], [29.0, 31.0]])
    t = 2.0 * np.pi * np.arange(6) / 6
    c1 = np.array([80.0, 20.0])
    ring1 = np.column_stack((c1[0] + 0.5 * np.cos(t), c1[1] + 0.5 * np.sin(t)))
    static = [(np.array([30.0, 30.0]), ring0), (c1, ring1)]

    loc[0, :2] = (22.0, 24.0); wp[0] = loc[0]
    speed[1] = 0.0
    loc[2, :2] = (3000.0, 3000.0); wp[2] = loc[2]; vel[2] = 0.0
    loc[3, :2] = (14.5, 11.0)                               # between (14, 10) and (15, 10)
    loc[4, :2] = (30.0, 30.0)                               # centre of the square ring
    loc[5, :2] = (200.0, 200.0)
    loc[6, :2] = (40.0, 40.0); loc[7, :2] = (42.0, 40.0)
    vel[6, :2] = (-0.75, 0.0); vel[7, :2] = (0.75, 0.0)     # lambda (v6 - v7) + e = (-3, 0) + (1, 0): antiparallel to d
    loc[8, :2] = (17.25, 10.5); mode[8] = synth.CROSSING_ROAD
    loc[9, :2] = (18.25, 9.5); mode[9] = ROAD_TO_SIDEWALK
    loc[10, :2] = centres[1] + (3.0, 4.0)                   # |.| = 5 exactly
    loc[11, :2] = c1 + (12.0, 16.0)                         # |.| = 20 exactly (static perception_threshold)
    loc[12, :2] = s0[7]
    loc[13, :2] = ring0[2]
    w = synth.Workload('appendix-b', 40.0, loc, vel, wp, radius, speed, mode, borders, centres, lengths, static,
                       veh_center=np.array([[25.0, 18.0], [35.0, 35.0]]), veh_yaw=np.array([0.0, 90.0]),
                       veh_vel=np.array([[5.0, 0.0], [0.0, 3.0]]), veh_extent=np.array([[2.4, 1.0], [2.4, 1.0]]))
    return w
