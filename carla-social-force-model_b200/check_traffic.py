"""Gap-acceptance test (call surface of the reference's ``check_traffic.py``) without the shapely dependency.

The reference intersects the pedestrian's straight path to its waypoint with each vehicle's straight path over the
crossing time using shapely LineStrings (check_traffic.py:30-54); two segments need no geometry library.  The result
follows shapely also where the geometry degenerates: collinear overlapping paths intersect in a LineString and the
three ``distance`` calls measure to its nearest point; a zero-length path is a point that either lies on the other
segment or does not.  K4a (``csrc/k4_lifecycle.cuh``) mirrors this function operation by operation.
"""
import numpy as np

from stateutils import normalize


def _segment_intersection(p0, p1, q0, q1):
    """Intersection of segments p0-p1 and q0-q1 as ``LineString.intersection`` gives it (check_traffic.py:46): None when
    empty, else (h0, h1) -- h0 == h1 for a point, the two ends of the overlap for collinear segments.  A pedestrian that
    stands on its waypoint (zero-length path) meets the vehicle's segment iff that point lies on it."""
    r, s = p1 - p0, q1 - q0
    denom = r[0] * s[1] - r[1] * s[0]
    qp = q0 - p0
    if denom != 0.0:
        t = (qp[0] * s[1] - qp[1] * s[0]) / denom
        u = (qp[0] * r[1] - qp[1] * r[0]) / denom
        if 0.0 <= t <= 1.0 and 0.0 <= u <= 1.0:
            h = p0 + t * r
            return h, h
        return None
    if qp[0] * r[1] - qp[1] * r[0] != 0.0:
        return None                                   # parallel, not collinear
    rr = r[0] * r[0] + r[1] * r[1]
    if rr == 0.0:
        ss = s[0] * s[0] + s[1] * s[1]
        if ss == 0.0:
            return (p0, p0) if (qp[0] == 0.0 and qp[1] == 0.0) else None
        if qp[0] * s[1] - qp[1] * s[0] != 0.0:
            return None
        t = -(qp[0] * s[0] + qp[1] * s[1]) / ss
        return (p0, p0) if 0.0 <= t <= 1.0 else None
    a = (qp[0] * r[0] + qp[1] * r[1]) / rr
    b = ((q1[0] - p0[0]) * r[0] + (q1[1] - p0[1]) * r[1]) / rr
    lo, hi = max(min(a, b), 0.0), min(max(a, b), 1.0)
    return (p0 + lo * r, p0 + hi * r) if lo <= hi else None


def _hit_distance(hit, x):
    """``intersection_point.distance(Point(x))`` (check_traffic.py:52-54): to the point, or to the nearest point of the
    overlap segment."""
    h0, h1 = hit
    v = h1 - h0
    vv = v[0] * v[0] + v[1] * v[1]
    if vv == 0.0:
        return np.linalg.norm(h0 - x)
    t = min(1.0, max(0.0, ((x[0] - h0[0]) * v[0] + (x[1] - h0[1]) * v[1]) / vv))
    return np.linalg.norm(h0 + t * v - x)


def check_traffic(ped, vehicles, vehicle_velocities, vehicle_extents):
    """True if the pedestrian can cross before / after every vehicle passes (check_traffic.py:7-61)."""
    ped_loc = np.asarray(ped['loc'][:2], dtype=float)
    ped_goal = np.asarray(ped['next_waypoint'][:2], dtype=float)
    ped_speed = ped['mode'].crossing_speed
    margin = ped['mode'].crossing_safety_margin
    if margin < 0:                                    # negative margin: cross without looking (:23-24)
        return True
    time_ped = np.linalg.norm(ped_goal - ped_loc) / ped_speed
    centres = np.array([c for c, _ in vehicles], dtype=float)
    velocities = np.asarray(vehicle_velocities, dtype=float)
    heading, _ = normalize(velocities)
    half_length = np.asarray(vehicle_extents)[:][0]   # sic: the reference indexes [:][0] (check_traffic.py:35-36)
    fronts = centres + heading * half_length
    backs = centres - heading * half_length
    for front, back, vel in zip(fronts, backs, velocities):
        veh_goal = front + vel * (time_ped + margin)
        hit = _segment_intersection(ped_loc, ped_goal, back, veh_goal)
        if hit is None:
            continue
        veh_speed = np.linalg.norm(vel)
        if veh_speed == 0:
            continue
        tti_ped = _hit_distance(hit, ped_loc) / ped_speed
        tti_front = _hit_distance(hit, front) / veh_speed
        tti_back = _hit_distance(hit, back) / veh_speed
        if tti_front - margin < tti_ped < tti_back + margin:
            return False
    return True
