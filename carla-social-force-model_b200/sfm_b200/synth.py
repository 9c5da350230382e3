"""Seeded synthetic crowds for the five BASELINE.json configurations.

Pure numpy, no device code.  Every coordinate is drawn in float32 and widened to float64, so the float64 oracle and the
float32 staging copies of the device path see bit-identical inputs.  The data contract matches what the reference's map
extraction hands to the forces (obstacles.py:269-281 ellipse rings, :332-359 straight borders with
``section_info = [middle point, n_points * resolution]``, run_simulation.py:192 ``[(centre, ring)]``).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

WALKING_SIDEWALK, CROSSING_ROAD = 1, 2          # ped_mode_manager.py:4-9


def _f32(a):
    return np.asarray(a, dtype=np.float32).astype(np.float64)


def ellipse_ring(centre, yaw_deg, extent_x, extent_y, resolution=0.1, size_factor=np.sqrt(2.0)):
    """Ring of border points around a box of half-extents (extent_x, extent_y), as obstacles.py:269-281 builds it.

    ``samples = max(6, int((2 ex + 2 ey) / resolution))`` points at equal parameter steps on the ellipse with
    semi-axes ``extent * size_factor``, rotated by the yaw and translated to the centre.
    """
    samples = max(6, int((2 * extent_x + 2 * extent_y) / resolution))
    t = 2.0 * np.pi * np.arange(samples) / samples
    px = extent_x * np.cos(t) * size_factor
    py = extent_y * np.sin(t) * size_factor
    c, s = np.cos(np.radians(yaw_deg)), np.sin(np.radians(yaw_deg))
    return np.column_stack((centre[0] + c * px - s * py, centre[1] + s * px + c * py))


def straight_border(start, heading, n_points, spacing):
    """Polyline of ``n_points`` at ``spacing`` metres from ``start`` along ``heading`` (rad) -- obstacles.py:349-351."""
    k = np.arange(n_points, dtype=np.float64)[:, None] * spacing
    return np.asarray(start, dtype=np.float64)[None, :] + k * np.array([np.cos(heading), np.sin(heading)])[None, :]


@dataclass
class Workload:
    name: str
    side: float
    loc: np.ndarray
    vel: np.ndarray
    next_waypoint: np.ndarray
    radius: np.ndarray
    target_speed: np.ndarray
    mode: np.ndarray
    borders: list = field(default_factory=list)
    section_center: np.ndarray = None
    section_length: np.ndarray = None
    static_obstacles: list = field(default_factory=list)
    veh_center: np.ndarray = None
    veh_yaw: np.ndarray = None
    veh_vel: np.ndarray = None
    veh_extent: np.ndarray = None
    veh_resolution: float = 0.1
    step_length: float = 0.05

    @property
    def n(self):
        return len(self.loc)

    def section_info(self):
        """(S, 2) object array [centre(2), length] -- the numpy>=1.24-safe form of forces.py:130-132's input."""
        info = np.empty((len(self.borders), 2), dtype=object)
        for s in range(len(self.borders)):
            info[s, 0] = self.section_center[s]
            info[s, 1] = float(self.section_length[s])
        return info

    def vehicles_at(self, step):
        """Dynamic-obstacle 6-tuple of pedestrian_simulation.py:108-115 at tick ``step`` (vehicles move ballistically)."""
        if self.veh_center is None or len(self.veh_center) == 0:
            return None
        centres = _f32(self.veh_center + self.veh_vel * (self.step_length * step))
        rings = [_f32(ellipse_ring(centres[v], self.veh_yaw[v], self.veh_extent[v, 0], self.veh_extent[v, 1],
                                   self.veh_resolution)) for v in range(len(centres))]
        ids = list(range(1000, 1000 + len(centres)))
        return (ids, [c for c in centres], list(self.veh_yaw), [v for v in self.veh_vel],
                [e for e in self.veh_extent], rings)


def make_crowd(n, side, rng, z_spread=0.0):
    loc = np.zeros((n, 3))
    loc[:, :2] = rng.uniform(0.0, side, size=(n, 2))
    vel = np.zeros((n, 3))
    vel[:, :2] = rng.normal(0.0, 1.0, size=(n, 2))
    if z_spread > 0.0:
        loc[:, 2] = rng.uniform(0.0, z_spread, size=n)
        vel[:, 2] = rng.normal(0.0, 0.05, size=n)
    wp = np.zeros((n, 3))
    wp[:, :2] = rng.uniform(0.0, side, size=(n, 2))
    radius = rng.uniform(0.2, 0.4, size=n)
    speed = rng.uniform(1.0, 1.6, size=n)
    mode = np.where(rng.random(n) < 0.9, WALKING_SIDEWALK, CROSSING_ROAD).astype(np.uint8)
    return _f32(loc), _f32(vel), _f32(wp), _f32(radius), _f32(speed), mode


def make_borders(n_sections, pts_per_section, spacing, length, side, rng):
    borders, centres = [], []
    for _ in range(n_sections):
        start = rng.uniform(0.0, side, size=2)
        line = _f32(straight_border(start, rng.uniform(0.0, 2 * np.pi), pts_per_section, spacing))
        borders.append(line)
        centres.append(line[len(line) // 2])                      # obstacles.py:352
    return borders, np.array(centres).reshape(-1, 2), np.full(n_sections, float(length))


def make_static_obstacles(n_obstacles, pts, ring_radius, side, rng):
    out = []
    t = 2.0 * np.pi * np.arange(pts) / pts
    for _ in range(n_obstacles):
        c = _f32(rng.uniform(0.0, side, size=2))
        ring = _f32(np.column_stack((c[0] + ring_radius * np.cos(t), c[1] + ring_radius * np.sin(t))))
        out.append((c, ring))
    return out


def make_vehicles(n_veh, side, rng, speed_max=15.0, extent=(2.4, 1.0)):
    centre = _f32(rng.uniform(0.0, side, size=(n_veh, 2)))
    yaw = _f32(rng.uniform(-180.0, 180.0, size=n_veh))
    speed = rng.uniform(0.0, speed_max, size=n_veh)
    vel = _f32(np.column_stack((speed * np.cos(np.radians(yaw)), speed * np.sin(np.radians(yaw)))))
    ext = np.tile(_f32(np.array(extent)), (n_veh, 1))
    return centre, yaw, vel, ext


# name, N, side, sections, pts/section, spacing, section_length, static obstacles, ring pts, vehicles
_CONFIGS = {
    1: ('cfg1', 64, 40.0, 5, 100, 0.4, 40.0, 10, 12, 2),
    2: ('cfg2', 4096, 64.0, 40, 200, 0.1, 20.0, 100, 12, 8),
    3: ('cfg3', 65536, 256.0, 5000, 210, 0.1, 20.0, 2500, 20, 0),
    4: ('cfg4', 262144, 512.0, 0, 0, 0.1, 20.0, 0, 20, 2048),
    5: ('cfg5', 1048576, 1024.0, 0, 0, 0.1, 20.0, 0, 20, 0),
}


def make_config(k, n=None, seed=None, z_spread=0.0, scale_sets=True):
    """Workload for BASELINE.json configs[k-1]; ``seed`` defaults to 1000+k.

    ``n`` overrides the pedestrian count keeping the density at 1 ped/m^2 (side = sqrt(n)); with ``scale_sets`` the
    border / obstacle / vehicle counts shrink in proportion to the area so a reduced case keeps the same neighbour
    statistics.
    """
    name, n0, side0, n_sec, pts, spacing, length, n_obs, ring_pts, n_veh = _CONFIGS[k]
    rng = np.random.default_rng(1000 + k if seed is None else seed)
    if n is None or n == n0:
        n, side, frac = n0, side0, 1.0
    else:
        side = float(np.float32(np.sqrt(n)))
        frac = (side * side) / (side0 * side0) if scale_sets else 1.0
        name = f'{name}-n{n}'
    n_sec, n_obs = int(round(n_sec * frac)), int(round(n_obs * frac))
    n_veh = int(round(n_veh * frac)) if k != 1 else n_veh
    loc, vel, wp, radius, speed, mode = make_crowd(n, side, rng, z_spread)
    w = Workload(name, side, loc, vel, wp, radius, speed, mode)
    if n_sec:
        w.borders, w.section_center, w.section_length = make_borders(n_sec, pts, spacing, length, side, rng)
    if n_obs:
        w.static_obstacles = make_static_obstacles(n_obs, ring_pts, 0.5, side, rng)
    if n_veh:
        if k == 1:      # two vehicles on fixed courses, 40-point rings
            w.veh_center = _f32(np.array([[5.0, 12.0], [28.0, 3.0]]))
            w.veh_yaw = _f32(np.array([0.0, 90.0]))
            w.veh_vel = _f32(np.array([[5.0, 0.0], [0.0, 3.0]]))
            w.veh_extent = np.tile(_f32(np.array([2.4, 1.0])), (2, 1))
            w.veh_resolution = 0.17
        else:
            w.veh_center, w.veh_yaw, w.veh_vel, w.veh_extent = make_vehicles(n_veh, side, rng)
    return w


# ---- lifecycle scenario: routes, mode machines, crossing traffic ---------------------------------
IDLE, ROAD_TO_SIDEWALK, CHECKING_TRAFFIC = 0, 3, 4


@dataclass
class Lifecycle:
    """What SimulationRunner / PedSpawner keep next to the PedState table (run_simulation.py:38-39,118-132,
    pedestrian_spawner.py:97-98,238-241): remaining waypoints per pedestrian and the PedModeManager constructor
    arguments.  ``idle`` pedestrians get ``set_mode(IDLE)`` at sim_time 0 before the first tick."""
    routes: list                      # per pedestrian: [(waypoint(3), crossing_road), ...]
    crossing_speed_factor: np.ndarray
    crossing_safety_margin: np.ndarray
    idle: np.ndarray                  # bool
    waypoint_threshold: float = 2.0
    despawn_on_arrival: bool = False
    spawn_tick: np.ndarray = None     # tick at whose start each pedestrian is spawned (None / 0: present from the start)


def make_lifecycle(n=48, seed=2001, n_vehicles=3, side=30.0, waypoints_per_ped=3, spawn_late=0, spawn_at=(20, 45)):
    """Small all-forces scene in which every lifecycle event happens within ~6 s: waypoint hand-overs (some requesting a
    road crossing), gap acceptance against crossing vehicles, idle pedestrians waking up, finished routes."""
    rng = np.random.default_rng(seed)
    loc, vel, wp, radius, speed, _ = make_crowd(n, side, rng)
    mode = np.full(n, WALKING_SIDEWALK, dtype=np.uint8)
    mode[rng.random(n) < 0.15] = CROSSING_ROAD
    mode[rng.random(n) < 0.15] = CHECKING_TRAFFIC
    # first waypoint 1.5 .. 4 m away so that hand-overs start early
    ang = rng.uniform(0.0, 2 * np.pi, size=n)
    wp[:, 0] = loc[:, 0] + rng.uniform(1.5, 4.0, size=n) * np.cos(ang)
    wp[:, 1] = loc[:, 1] + rng.uniform(1.5, 4.0, size=n) * np.sin(ang)
    wp = _f32(wp)
    routes = []
    for i in range(n):
        cur, route = wp[i].copy(), []
        for _ in range(int(rng.integers(0, waypoints_per_ped + 1))):
            a = rng.uniform(0.0, 2 * np.pi)
            cur = _f32(cur + np.array([3.0 * np.cos(a), 3.0 * np.sin(a), 0.0]))
            route.append((cur.copy(), bool(rng.random() < 0.5)))
        routes.append(route)
    w = Workload(f'lifecycle-n{n}', side, loc, vel, wp, radius, speed, mode)
    w.borders, w.section_center, w.section_length = make_borders(4, 100, 0.4, 40.0, side, rng)
    w.static_obstacles = make_static_obstacles(6, 12, 0.5, side, rng)
    w.veh_center = _f32(rng.uniform(0.0, side, size=(n_vehicles, 2)))
    w.veh_yaw = _f32(rng.uniform(-180.0, 180.0, size=n_vehicles))
    sp = rng.uniform(2.0, 9.0, size=n_vehicles)
    w.veh_vel = _f32(np.column_stack((sp * np.cos(np.radians(w.veh_yaw)), sp * np.sin(np.radians(w.veh_yaw)))))
    w.veh_vel[0] = 0.0                                     # a parked vehicle: the zero-speed branch of check_traffic.py:48
    w.veh_extent = np.tile(_f32(np.array([2.4, 1.0])), (n_vehicles, 1))
    w.veh_resolution = 0.17
    margin = _f32(rng.uniform(0.5, 2.5, size=n))
    margin[rng.random(n) < 0.1] = -1.0                      # crosses without looking (check_traffic.py:24)
    life = Lifecycle(routes, _f32(rng.uniform(1.2, 1.8, size=n)), margin, rng.random(n) < 0.12)
    if spawn_late:                                          # the last `spawn_late` pedestrians join in two waves
        tick = np.zeros(n, dtype=np.int64)
        tick[n - spawn_late:n - spawn_late // 2] = spawn_at[0]
        tick[n - spawn_late // 2:] = spawn_at[1]
        life.spawn_tick = tick
    return w, life


def make_output_scene(n=5, frames=4, seed=3001):
    """A tiny recorded run in the containers ``OutputGenerator`` reads (output_generator.py:12-16): per-tick pedestrian
    snapshots (pedestrian_state.py:100-104), per-tick vehicle snapshots (pedestrian_simulation.py:129-140), static
    obstacles ``[(centre, ring)]`` and border polylines.  Values include sub-1e-4 and > 1e16 magnitudes so that the number
    formatting (``str`` of a float64) is exercised on both sides of its notation switches."""
    rng = np.random.default_rng(seed)
    dtype = [('name', 'U8'), ('id', 'i4'), ('loc', 'f8', (3,)), ('vel', 'f8', (3,)), ('next_waypoint', 'f8', (3,)),
             ('mode', 'O'), ('radius', 'f8'), ('target_speed', 'f8')]
    vdtype = [('id', 'i4'), ('loc', 'f8', (2,)), ('heading', 'f8'), ('vel', 'f8', (2,)), ('extent', 'f8', (2,))]
    ped_states, veh_states = {}, {}
    for k in range(frames):
        st = np.zeros(n, dtype=dtype)
        st['name'] = [f'ped_{i + 3}' for i in range(n)]
        st['id'] = np.arange(n) + 100
        st['loc'] = rng.normal(0.0, 30.0, size=(n, 3))
        st['vel'] = rng.normal(0.0, 1.0, size=(n, 3))
        st['loc'][0, 0], st['vel'][0, 1] = 1.25e-7 * (k + 1), 3.0e17
        st['mode'] = [int(m) for m in rng.integers(0, 5, size=n)]
        ped_states[k * 0.05] = st
        vs = np.zeros(2, dtype=vdtype)
        vs['id'] = [1000, 1001]
        vs['loc'] = rng.uniform(0.0, 40.0, size=(2, 2))
        vs['heading'] = rng.uniform(-180.0, 180.0, size=2)
        vs['vel'] = rng.normal(0.0, 5.0, size=(2, 2))
        vs['extent'] = [[2.4, 1.0], [2.4, 1.0]]
        veh_states[k * 0.05] = vs
    static = make_static_obstacles(2, 6, 0.5, 40.0, rng)
    borders, _, _ = make_borders(2, 5, 0.4, 40.0, 40.0, rng)
    return dict(ped_states=ped_states, veh_states=veh_states, static_obstacles=static, borders=borders)
