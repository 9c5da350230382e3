"""Import the *unmodified* reference (``/root/reference``) so the oracle can be pinned against it.

TEST INFRASTRUCTURE.  Works only in the build container: the GPU box has no ``/root/reference``, so nothing that runs
there imports this module (``available()`` is the guard).  Two shims, both documented in SURVEY.md section 8c:

* ``shapely`` (check_traffic.py:2) is not installed -> a stand-in module is registered before the import with the two
  classes ``check_traffic`` uses (``LineString([a, b]).intersection(other)`` -> something with ``is_empty`` and
  ``distance(Point)``), restricted to two-point segments.  The reference's own gap-acceptance code then runs unmodified;
  only the intersection primitive is ours (parity unpinned for shapely itself).
* ``BorderForce.__init__`` does ``np.array(section_info)`` on a ragged list (forces.py:130), which numpy >= 1.24
  rejects -> ``section_info`` is passed as a pre-built ``dtype=object`` array, the form the reference's own ``.npz`` cache
  yields (obstacles.py:43-45).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

REFERENCE_DIR = os.environ.get('SFM_REFERENCE_DIR', '/root/reference')
_OWN_MODULES = ('forces', 'stateutils', 'pedestrian_state', 'pedestrian_simulation', 'ped_mode_manager', 'check_traffic')


class _Point:
    """Stand-in for shapely.geometry.Point (2-D)."""
    is_empty = False

    def __init__(self, xy):
        self.xy = np.asarray(xy, dtype=np.float64)[:2]

    def distance(self, other):
        return float(np.linalg.norm(self.xy - other.xy))


class _Empty:
    is_empty = True


class _Overlap:
    """What shapely returns for two collinear overlapping segments: the LineString of the overlap; ``distance`` to a
    point is the distance to its nearest point."""
    is_empty = False

    def __init__(self, h0, h1):
        self.hit = (h0, h1)

    def distance(self, other):
        from oracle.lifecycle_oracle import hit_distance
        return float(hit_distance(self.hit, other.xy))


class _LineString:
    """Stand-in for shapely.geometry.LineString restricted to one segment."""

    def __init__(self, coords):
        self.a, self.b = (np.asarray(c, dtype=np.float64)[:2] for c in coords)

    def intersection(self, other):
        from oracle.lifecycle_oracle import segment_intersection
        hit = segment_intersection(self.a, self.b, other.a, other.b)
        if hit is None:
            return _Empty()
        return _Point(hit[0]) if hit[0] is hit[1] or np.array_equal(hit[0], hit[1]) else _Overlap(*hit)


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, 'forces.py'))


def load():
    """Return a namespace with the reference's modules, imported under private names.

    The reference's modules import each other by bare top-level name (forces.py:7-8), and the drop-in package uses the
    same names, so the import happens with ``REFERENCE_DIR`` first on ``sys.path`` and the resulting modules are then
    *removed* from ``sys.modules`` again -- both implementations can live in one test process.
    """
    if not available():
        raise RuntimeError(f'reference not found at {REFERENCE_DIR}')
    saved = {name: sys.modules.pop(name) for name in _OWN_MODULES if name in sys.modules}
    saved_flag = sys.dont_write_bytecode
    sys.dont_write_bytecode = True                       # the reference tree is read-only
    if 'shapely' not in sys.modules:
        sh, geo = types.ModuleType('shapely'), types.ModuleType('shapely.geometry')
        geo.LineString, geo.Point = _LineString, _Point
        sh.geometry = geo
        sys.modules['shapely'], sys.modules['shapely.geometry'] = sh, geo
    sys.path.insert(0, REFERENCE_DIR)
    try:
        import forces, stateutils, pedestrian_state, pedestrian_simulation, ped_mode_manager   # noqa: E401
        ns = types.SimpleNamespace(forces=forces, stateutils=stateutils, pedestrian_state=pedestrian_state,
                                   pedestrian_simulation=pedestrian_simulation, ped_mode_manager=ped_mode_manager)
    finally:
        sys.path.remove(REFERENCE_DIR)
        for name in _OWN_MODULES:
            sys.modules.pop(name, None)
        sys.modules.update(saved)
        sys.dont_write_bytecode = saved_flag
    return ns


def load_config():
    import tomllib
    with open(os.path.join(REFERENCE_DIR, 'config', 'sfm_config.toml'), 'rb') as f:
        return tomllib.load(f)


def build_simulation(ref, workload, sfm_config):
    """A reference ``PedestrianSimulation`` holding the workload's pedestrians (bulk-filled structured array)."""
    sim = ref.pedestrian_simulation.PedestrianSimulation(list(workload.borders), workload.section_info(),
                                                         list(workload.static_obstacles), sfm_config,
                                                         workload.step_length)
    PedMode, Manager = ref.ped_mode_manager.PedMode, ref.ped_mode_manager.PedModeManager
    state = np.zeros(workload.n, dtype=sim.peds.ped_state_dtype)
    state['name'] = [f'p{i}'[:8] for i in range(workload.n)]
    state['id'] = np.arange(workload.n)
    state['loc'], state['vel'], state['next_waypoint'] = workload.loc, workload.vel, workload.next_waypoint
    state['radius'], state['target_speed'] = workload.radius, workload.target_speed
    for i in range(workload.n):
        # crossing_speed_factor 1.0 keeps target_speed == the drawn value in both modes (ped_mode_manager.py:22,61-63)
        m = Manager(state['name'][i], float(workload.target_speed[i]), PedMode(int(workload.mode[i])), 1.0, -1.0)
        state['mode'][i] = m
    sim.peds.state = state
    return sim


def run_ticks(ref, workload, sfm_config, n_steps, record_forces=True):
    """Drive the reference's own tick loop headless: tick -> read new velocities -> x += dt * v (the CARLA stub).

    Returns dict(loc=(T+1,N,3), vel=(T+1,N,3), forces={class: (T,N,3)}).
    """
    sim = build_simulation(ref, workload, sfm_config)
    dt = workload.step_length
    locs, vels = [sim.peds.state['loc'].copy()], [sim.peds.state['vel'].copy()]
    forces = {name: [] for name in sim.forces}
    for step in range(n_steps):
        veh = workload.vehicles_at(step)
        if veh is not None:
            sim.update_dynamic_obstacles(veh)                                    # pedestrian_simulation.py:108-115
        if record_forces:
            for name, f in sim.forces.items():
                forces[name].append(f.get_force(sim.peds))
        sim.tick(step * dt)                                                      # pedestrian_simulation.py:57-83
        nv = sim.get_new_velocities()                                            # aliases state['vel']
        sim.peds.state['loc'] += nv['vel'] * dt                                  # CARLA stub
        sim.peds.all_states.clear()                                              # unbounded history, not needed
        locs.append(sim.peds.state['loc'].copy())
        vels.append(sim.peds.state['vel'].copy())
    return dict(loc=np.array(locs), vel=np.array(vels), forces={k: np.array(v) for k, v in forces.items()})


def run_lifecycle(ref, workload, life, sfm_config, n_steps, despawn=False):
    """The reference's tick loop with its own mode machines, gap acceptance and waypoint hand-over, CARLA stubbed.

    Follows SimulationRunner.tick (run_simulation.py:94-132) for everything that does not need the simulator: vehicles
    refreshed -> ``PedestrianSimulation.tick`` -> arrival test + ``update_next_waypoint`` -> positions advanced by the
    stub.  Pedestrians are spawned through ``spawn_pedestrian`` with real ``PedModeManager`` objects
    (pedestrian_spawner.py:238-241)."""
    sim = ref.pedestrian_simulation.PedestrianSimulation(list(workload.borders), workload.section_info(),
                                                         list(workload.static_obstacles), sfm_config,
                                                         workload.step_length)
    PedMode, Manager = ref.ped_mode_manager.PedMode, ref.ped_mode_manager.PedModeManager
    n = workload.n
    names = [f'p_{i}' for i in range(n)]
    spawn_tick = getattr(life, 'spawn_tick', None)
    spawn_tick = np.zeros(n, dtype=np.int64) if spawn_tick is None else np.asarray(spawn_tick)

    def spawn(i):
        m = Manager(names[i], float(workload.target_speed[i]), PedMode(int(workload.mode[i])),
                    float(life.crossing_speed_factor[i]), float(life.crossing_safety_margin[i]))
        if life.idle[i]:
            m.set_mode(PedMode.IDLE)
        sim.spawn_pedestrian((names[i], i, workload.loc[i], workload.vel[i], workload.next_waypoint[i], m,
                              float(workload.radius[i]), float(workload.target_speed[i])))

    for i in np.nonzero(spawn_tick == 0)[0]:
        spawn(i)
    waypoint_dict = {names[i]: list(life.routes[i]) for i in range(n)}
    dt = workload.step_length
    codes = lambda: np.array([int(m.current_mode) for m in sim.peds.state['mode']], dtype=np.uint8)   # noqa: E731
    speeds = lambda: np.array([float(m.target_speed) for m in sim.peds.state['mode']])                # noqa: E731
    hist = dict(loc=[sim.peds.state['loc'].copy()], vel=[sim.peds.state['vel'].copy()], mode=[codes()],
                wp=[sim.peds.state['next_waypoint'].copy()], target_speed=[], mode_speed=[speeds()],
                remaining=[np.array([len(waypoint_dict[k]) for k in sim.peds.state['name']])])
    ragged = despawn or bool(spawn_tick.any())
    if ragged:
        hist['ids'] = [sim.peds.state['id'].copy()]
    for step in range(n_steps):
        if step > 0:
            for i in np.nonzero(spawn_tick == step)[0]:                        # the spawner runs before the tick
                spawn(i)
        veh = workload.vehicles_at(step)
        if veh is not None:
            sim.update_dynamic_obstacles(veh)
        sim.tick(step * dt)
        nv = sim.get_new_velocities()
        hist['target_speed'].append(sim.peds.state['target_speed'].copy())
        for ped_name in sim.get_arrived_peds(life.waypoint_threshold):        # run_simulation.py:118-125
            remaining = waypoint_dict[ped_name]
            if remaining:
                sim.peds.update_next_waypoint(ped_name, remaining.pop(0))
            elif despawn:                                                      # run_simulation.py:127-132
                sim.destroy_pedestrian(ped_name)
                waypoint_dict.pop(ped_name)
        if ragged:
            if sim.peds.size() == 0:
                break
            nv = sim.peds.state[['id', 'vel']]
            hist['ids'].append(sim.peds.state['id'].copy())
        sim.peds.state['loc'] += nv['vel'] * dt
        sim.peds.all_states.clear()
        sim.all_dyn_obs_states.clear()
        hist['loc'].append(sim.peds.state['loc'].copy()); hist['vel'].append(sim.peds.state['vel'].copy())
        hist['mode'].append(codes()); hist['wp'].append(sim.peds.state['next_waypoint'].copy())
        hist['mode_speed'].append(speeds())
        hist['remaining'].append(np.array([len(waypoint_dict[k]) for k in sim.peds.state['name']]))
    if ragged:
        return hist
    return {k: np.array(v) for k, v in hist.items()}
