"""Hand-built gap-acceptance scenes in which the pedestrian's path and a vehicle's path are COLLINEAR or degenerate -- the
cases in which shapely's ``LineString.intersection`` (check_traffic.py:46) is not a point: an overlap segment whose
``distance`` to the three reference points is measured to its nearest point, or a zero-length path.

Each scene: (ped_loc, ped_goal, crossing_speed, safety_margin, vehicle centres, velocities, extents, expected decision).
The expected decisions were worked out by hand from check_traffic.py:27-58 with shapely's semantics (see the comments).
"""
import numpy as np

EXT = np.array([[2.4, 1.0]])


def scenes():
    out = []
    # 1. head-on along the x axis: back (22.4, 0) -> goal (-13, 0) covers the whole crossing [0, 10]; the overlap touches the
    #    pedestrian (tti_ped = 0), the front is 7.6 m from it (1.52 s): 1.52 - 1 < 0 fails -> free to cross.  (Measured to
    #    the overlap's midpoint instead, the same scene is blocked: 1.52 < 2.56 < 4.48.)
    out.append(((0.0, 0.0), (10.0, 0.0), 1.95, 1.0, [[20.0, 0.0]], [[-5.0, 0.0]], EXT, True))
    # 2. the same vehicle, already over the start of the crossing: front (-0.4, 0), back (4.4, 0): the overlap [0, 4.4]
    #    contains the pedestrian and the back (distance 0 both), the front is 0.4 m off: 0.08 - 1 < 0 < 0 + 1 -> blocked
    out.append(((0.0, 0.0), (10.0, 0.0), 1.95, 1.0, [[2.0, 0.0]], [[-5.0, 0.0]], EXT, False))
    # 3. vehicle behind the pedestrian driving the same way, overlap starts at the pedestrian: back (-12.4, 0) -> goal far
    #    ahead; tti_ped = 0, tti_front = 7.6 / 5, tti_back = 12.4 / 5: 0.52 < 0 fails -> free
    out.append(((0.0, 0.0), (10.0, 0.0), 1.95, 1.0, [[-10.0, 0.0]], [[5.0, 0.0]], EXT, True))
    # 4. collinear but disjoint (vehicle drives away beyond the goal): empty intersection -> free
    out.append(((0.0, 0.0), (10.0, 0.0), 1.95, 1.0, [[30.0, 0.0]], [[5.0, 0.0]], EXT, True))
    # 5. oblique collinear pair (direction (3, 4) / 5), vehicle over the second half of the crossing
    d = np.array([0.6, 0.8])
    out.append((tuple(0.0 * d), tuple(10.0 * d), 1.95, 0.5, [list(8.0 * d)], [list(-4.0 * d)], EXT, None))
    # 6. the pedestrian stands on its waypoint (zero-length path) in the lane of a passing vehicle: the point lies on the
    #    vehicle's segment; time_ped = 0, tti_ped = 0, front 7.6 m away at 5 m/s: 1.52 - 2 < 0 < tti_back + 2 -> blocked
    out.append(((5.0, 0.0), (5.0, 0.0), 1.95, 2.0, [[-5.0, 0.0]], [[5.0, 0.0]], EXT, False))
    # 7. the same pedestrian beside the lane: the point is not on the segment -> free
    out.append(((5.0, 3.0), (5.0, 3.0), 1.95, 2.0, [[-5.0, 0.0]], [[5.0, 0.0]], EXT, True))
    # 8. parallel, not collinear -> free;  9. a parked vehicle on the path (speed 0 is skipped, check_traffic.py:48) -> free
    out.append(((0.0, 0.0), (10.0, 0.0), 1.95, 1.0, [[20.0, 1.5]], [[-5.0, 0.0]], EXT, True))
    out.append(((0.0, 0.0), (10.0, 0.0), 1.95, 1.0, [[5.0, 0.0]], [[0.0, 0.0]], EXT, True))
    return [(np.array(a, dtype=float), np.array(b, dtype=float), s, m, np.array(c, dtype=float), np.array(v, dtype=float),
             np.array(e, dtype=float), want) for a, b, s, m, c, v, e, want in out]
