"""CPU oracle for the per-tick bookkeeping around the forces  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Array-form numpy restatement of what the reference does in interpreter loops around the force sum (SURVEY.md section
8f): the pedestrian mode machines (ped_mode_manager.py:12-70), the gap-acceptance test (check_traffic.py:7-61, with the
two shapely LineStrings replaced by a closed-form segment intersection), the arrival test and waypoint hand-over
(pedestrian_simulation.py:88-97, run_simulation.py:118-132, pedestrian_state.py:83-95) and the vehicle ellipse rings
(obstacles.py:269-281).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import it.

Parity pin: ``tests/test_lifecycle_oracle.py`` drives the *imported reference* (its own PedModeManager objects, its own
``PedestrianSimulation.tick`` and its own ``check_traffic`` running against a minimal functional stand-in for the two
shapely classes it uses) through the same scenario and asserts identical mode sequences and waypoints; the committed
``tests/golden/lifecycle.npz`` (``oracle/make_golden.py --only lifecycle``) carries that run to the GPU box.
*Parity unpinned*: shapely itself (``Shapely==1.6.4.post2``, requirements.txt:5) is not installed, so the segment
intersection primitive is ours on both sides; and ``carla.Transform.transform`` (float32 inside the CARLA client
library) is restated in float64 for the vehicle rings.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from oracle import sfm_oracle as O

IDLE, WALKING_SIDEWALK, CROSSING_ROAD, ROAD_TO_SIDEWALK, CHECKING_TRAFFIC = 0, 1, 2, 3, 4


@dataclass
class Machines:
    """The fields of N PedModeManager objects as columns (ped_mode_manager.py:18-28)."""
    mode: np.ndarray                  # current_mode, uint8
    target_speed: np.ndarray          # mode.target_speed
    initial_target_speed: np.ndarray
    crossing_speed: np.ndarray
    crossing_safety_margin: np.ndarray
    next_mode_time: np.ndarray
    waiting_time: float = 5.0
    sim_time: float = 0.0

    @classmethod
    def create(cls, target_speed, initial_mode, crossing_speed_factor, crossing_safety_margin):
        ts = np.asarray(target_speed, dtype=np.float64)
        return cls(np.array(initial_mode, dtype=np.uint8), ts.copy(), ts.copy(),
                   np.asarray(crossing_speed_factor, dtype=np.float64) * ts,                  # :22
                   np.asarray(crossing_safety_margin, dtype=np.float64).copy(), np.full(len(ts), -1.0))

    def copy(self):
        return Machines(self.mode.copy(), self.target_speed.copy(), self.initial_target_speed.copy(),
                        self.crossing_speed.copy(), self.crossing_safety_margin.copy(), self.next_mode_time.copy(),
                        self.waiting_time, self.sim_time)

    # ped_mode_manager.py:49-69
    def activate(self, rows, mode):
        rows = np.asarray(rows)
        if mode == IDLE:
            self.target_speed[rows] = 0.0
            self.next_mode_time[rows] = self.sim_time + self.waiting_time
        elif mode == WALKING_SIDEWALK:
            self.target_speed[rows] = self.initial_target_speed[rows]
        elif mode == CROSSING_ROAD:
            self.target_speed[rows] = self.crossing_speed[rows]
        elif mode == CHECKING_TRAFFIC:
            self.target_speed[rows] = 0.0
        elif mode != ROAD_TO_SIDEWALK:
            return
        self.mode[rows] = mode

    # ped_mode_manager.py:37-47
    def set_mode(self, rows, wanted):
        rows = np.atleast_1d(np.asarray(rows))
        cur = self.mode[rows]
        if wanted == CROSSING_ROAD:
            detour = cur == WALKING_SIDEWALK
            self.activate(rows[detour], CHECKING_TRAFFIC)
            self.activate(rows[~detour], CROSSING_ROAD)
        elif wanted == WALKING_SIDEWALK:
            detour = cur == CROSSING_ROAD
            self.activate(rows[detour], ROAD_TO_SIDEWALK)
            self.activate(rows[~detour], WALKING_SIDEWALK)
        else:
            self.activate(rows, wanted)

    # ped_mode_manager.py:30-35
    def tick(self, sim_time):
        self.sim_time = sim_time
        wake = (self.mode == IDLE) & (self.next_mode_time <= sim_time)
        self.activate(np.nonzero(wake)[0], WALKING_SIDEWALK)


def segment_intersection(p0, p1, q0, q1):
    """Intersection of segments p0-p1 and q0-q1 as shapely's ``LineString.intersection`` yields it (check_traffic.py:46):
    None when empty, else the pair (h0, h1) -- a Point has h0 == h1, a collinear overlap is the LineString h0-h1.  A
    zero-length pedestrian path (the pedestrian stands on its waypoint) meets the other segment iff that point lies on it."""
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
        return None
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


def hit_distance(hit, x):
    """``intersection.distance(Point(x))`` (check_traffic.py:52-54): distance from x to the point, or to the nearest point
    of the overlap segment."""
    h0, h1 = hit
    v = h1 - h0
    vv = v[0] * v[0] + v[1] * v[1]
    if vv == 0.0:
        return np.linalg.norm(h0 - x)
    t = min(1.0, max(0.0, ((x[0] - h0[0]) * v[0] + (x[1] - h0[1]) * v[1]) / vv))
    return np.linalg.norm(h0 + t * v - x)


def check_traffic(ped_loc, ped_goal, crossing_speed, safety_margin, veh_centres, veh_velocities, veh_extents):
    """check_traffic.py:7-61 for one pedestrian.  ``veh_extents[:][0]`` -- the first vehicle's extent -- is what the
    reference multiplies every heading with (:35-36); that is kept."""
    if safety_margin < 0:                                                     # :24
        return True
    ped_loc, ped_goal = np.asarray(ped_loc[:2], dtype=np.float64), np.asarray(ped_goal[:2], dtype=np.float64)
    time_ped = np.linalg.norm(ped_goal - ped_loc) / crossing_speed              # :27-28
    centres = np.asarray(veh_centres, dtype=np.float64).reshape(-1, 2)
    velocities = np.asarray(veh_velocities, dtype=np.float64).reshape(-1, 2)
    heading, _ = O.normalize(velocities)                                      # :34
    half = np.asarray(veh_extents, dtype=np.float64).reshape(-1, 2)[0]        # :35-36 (sic)
    fronts, backs = centres + heading * half, centres - heading * half
    for front, back, vel in zip(fronts, backs, velocities):
        goal = front + vel * (time_ped + safety_margin)                       # :42
        hit = segment_intersection(ped_loc, ped_goal, back, goal)             # :43-46
        if hit is None:
            continue
        speed = np.linalg.norm(vel)
        if speed == 0:                                                        # :48-49
            continue
        tti_ped = hit_distance(hit, ped_loc) / crossing_speed
        tti_front = hit_distance(hit, front) / speed
        tti_back = hit_distance(hit, back) / speed
        if tti_front - safety_margin < tti_ped < tti_back + safety_margin:    # :57
            return False
    return True


def tick_modes(machines, state_target_speed, loc, next_waypoint, sim_time, vehicles=None):
    """The bookkeeping half of PedestrianSimulation.tick (pedestrian_simulation.py:63-73), in place.

    ``vehicles`` = (centres, velocities, extents) or None.  Returns the number of pedestrians that started crossing."""
    state_target_speed[:] = machines.target_speed                              # pedestrian_state.py:94-95
    machines.tick(sim_time)                                                    # :64-65
    started = 0
    for i in np.nonzero(machines.mode == CHECKING_TRAFFIC)[0]:                 # :67
        ready = True
        if vehicles is not None and len(vehicles[0]):
            ready = check_traffic(loc[i], next_waypoint[i], machines.crossing_speed[i],
                                  machines.crossing_safety_margin[i], *vehicles)
        if ready:
            machines.set_mode(i, CROSSING_ROAD)                                # :73
            started += 1
    return started


def advance_waypoints(machines, loc, next_waypoint, routes, cursor, finished, threshold):
    """get_arrived_peds (pedestrian_simulation.py:88-97) + the hand-over loop of SimulationRunner.tick
    (run_simulation.py:118-125) + PedState.update_next_waypoint (pedestrian_state.py:83-92), in place.
    ``routes[i]`` is the pedestrian's full list of (waypoint(3), crossing_road); ``cursor[i]`` the next unread entry."""
    diff = next_waypoint[:, :2] - loc[:, :2]
    arrived = np.nonzero(np.linalg.norm(diff, axis=-1) < threshold)[0]
    for i in arrived:
        if cursor[i] < len(routes[i]):
            wp, crossing = routes[i][cursor[i]]
            cursor[i] += 1
            next_waypoint[i] = wp
            machines.set_mode(i, CROSSING_ROAD if crossing else WALKING_SIDEWALK)
        else:
            finished[i] = True
    return arrived


def ellipse_ring(centre, yaw_deg, extent_x, extent_y, resolution=0.1, size_factor=np.sqrt(2.0)):
    """obstacles.py:269-281 with carla.Transform.transform restated for pitch = roll = 0 (float64)."""
    circumference = 2 * extent_x + 2 * extent_y
    samples = max([6, int(circumference / resolution)])
    out = np.empty((samples, 2))
    yaw = np.radians(yaw_deg)
    cy, sy = np.cos(yaw), np.sin(yaw)
    for i in range(samples):
        theta = np.pi * 2 * i / samples
        x, y = extent_x * np.cos(theta) * size_factor, extent_y * np.sin(theta) * size_factor
        out[i] = (centre[0] + (cy * x - sy * y), centre[1] + (sy * x + cy * y))
    return out


def run_headless(scene, w, life, n_steps, vehicles_at=None, despawn=False):
    """The whole per-tick loop with CARLA stubbed: vehicles -> mode machines + gap acceptance -> forces, velocity
    update -> arrival test at the positions the forces saw -> x += dt v.

    With ``despawn`` pedestrians that arrive with no waypoint left are removed right after the hand-over loop
    (run_simulation.py:127-132); with ``life.spawn_tick`` pedestrians join the crowd at the start of their tick
    (pedestrian_simulation.py:99-100).  Histories then carry ``ids`` (original row of every present pedestrian) per tick
    and ragged state lists.

    Returns per-tick histories (T+1 entries for state, T for decisions)."""
    n = w.n
    spawn_tick = getattr(life, 'spawn_tick', None)
    spawn_tick = np.zeros(n, dtype=np.int64) if spawn_tick is None else np.asarray(spawn_tick)
    ragged = despawn or bool(spawn_tick.any())

    def rows_of(sel):
        """State columns of the pedestrians ``sel`` as the spawner hands them over (pedestrian_spawner.py:238-241)."""
        m = Machines.create(w.target_speed[sel], w.mode[sel], np.asarray(life.crossing_speed_factor)[sel],
                            np.asarray(life.crossing_safety_margin)[sel])
        m.set_mode(np.nonzero(np.asarray(life.idle)[sel])[0], IDLE)
        return m, w.loc[sel].copy(), w.vel[sel].copy(), w.next_waypoint[sel].copy(), w.target_speed[sel].copy(), \
            w.radius[sel].copy(), [life.routes[i] for i in sel]

    ids = np.nonzero(spawn_tick == 0)[0]
    machines, loc, vel, wp, target_speed, radius, routes = rows_of(ids)
    cursor, finished = np.zeros(len(ids), dtype=np.int64), np.zeros(len(ids), dtype=bool)
    hist = dict(ids=[ids.copy()], loc=[loc.copy()], vel=[vel.copy()], mode=[machines.mode.copy()], wp=[wp.copy()],
                target_speed=[], mode_speed=[machines.target_speed.copy()], cursor=[cursor.copy()],
                finished=[finished.copy()])
    vehicles_at = vehicles_at or w.vehicles_at
    for step in range(n_steps):
        t = step * w.step_length
        late = np.nonzero(spawn_tick == step)[0] if step > 0 else np.zeros(0, dtype=np.int64)
        if len(late):                                   # spawned at the start of the tick, appended in index order
            m2, loc2, vel2, wp2, ts2, rad2, routes2 = rows_of(late)
            m2.sim_time = machines.sim_time
            machines = Machines(*(np.concatenate((getattr(machines, f), getattr(m2, f))) for f in
                                  ('mode', 'target_speed', 'initial_target_speed', 'crossing_speed',
                                   'crossing_safety_margin', 'next_mode_time')), machines.waiting_time, machines.sim_time)
            loc, vel, wp = np.concatenate((loc, loc2)), np.concatenate((vel, vel2)), np.concatenate((wp, wp2))
            target_speed, radius = np.concatenate((target_speed, ts2)), np.concatenate((radius, rad2))
            routes = routes + routes2
            cursor = np.concatenate((cursor, np.zeros(len(late), dtype=np.int64)))
            finished = np.concatenate((finished, np.zeros(len(late), dtype=bool)))
            ids = np.concatenate((ids, late))
        veh = vehicles_at(step)
        traffic = (veh[1], veh[3], veh[4]) if veh is not None else None
        tick_modes(machines, target_speed, loc, wp, t, traffic)
        dyn = list(zip(veh[1], veh[5])) if veh is not None else None
        new_loc, new_vel, _ = O.step(scene, loc, vel, wp, radius, target_speed, machines.mode, dyn,
                                     veh[3] if veh is not None else None)
        advance_waypoints(machines, loc, wp, routes, cursor, finished, life.waypoint_threshold)
        loc, vel = new_loc, new_vel
        if despawn and finished.any():
            keep = ~finished
            loc, vel, wp, radius, target_speed = loc[keep], vel[keep], wp[keep], radius[keep], target_speed[keep]
            cursor, ids = cursor[keep], ids[keep]
            routes = [r for r, k in zip(routes, keep) if k]
            machines = Machines(machines.mode[keep], machines.target_speed[keep], machines.initial_target_speed[keep],
                                machines.crossing_speed[keep], machines.crossing_safety_margin[keep],
                                machines.next_mode_time[keep], machines.waiting_time, machines.sim_time)
            finished = np.zeros(len(ids), dtype=bool)
        hist['ids'].append(ids.copy())
        hist['loc'].append(loc.copy()); hist['vel'].append(vel.copy()); hist['mode'].append(machines.mode.copy())
        hist['wp'].append(wp.copy()); hist['target_speed'].append(target_speed.copy())
        hist['mode_speed'].append(machines.target_speed.copy()); hist['cursor'].append(cursor.copy())
        hist['finished'].append(finished.copy())
    if ragged:
        return hist
    return {k: np.array(v) for k, v in hist.items()}
