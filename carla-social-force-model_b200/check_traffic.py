"""Gap-acceptance test (call surface of the reference's ``check_traffic.py``) without the shapely dependency.

The reference intersects the pedestrian's straight path to its waypoint with each vehicle's straight path over the
crossing time using shapely LineStrings (check_traffic.py:30-54); two segments need no geometry library.
"""
import numpy as np

from stateutils import normalize


def _segment_intersection(p0, p1, q0, q1):
    """Intersection of segments p0-p1 and q0-q1: None, or the point (for collinear overlap: the overlap's midpoint)."""
    r, s = p1 - p0, q1 - q0
    denom = r[0] * s[1] - r[1] * s[0]
    qp = q0 - p0
    if denom != 0.0:
        t = (qp[0] * s[1] - qp[1] * s[0]) / denom
        u = (qp[0] * r[1] - qp[1] * r[0]) / denom
        return p0 + t * r if (0.0 <= t <= 1.0 and 0.0 <= u <= 1.0) else None
    if qp[0] * r[1] - qp[1] * r[0] != 0.0:
        return None                                   # parallel, not collinear
    rr = float(r @ r)
    if rr == 0.0:
        return None
    t0, t1 = sorted((float(qp @ r) / rr, float((q1 - p0) @ r) / rr))
    lo, hi = max(t0, 0.0), min(t1, 1.0)
    return p0 + 0.5 * (lo + hi) * r if lo <= hi else None


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
        tti_ped = np.linalg.norm(hit - ped_loc) / ped_speed
        tti_front = np.linalg.norm(hit - front) / veh_speed
        tti_back = np.linalg.norm(hit - back) / veh_speed
        if tti_front - margin < tti_ped < tti_back + margin:
            return False
    return True
