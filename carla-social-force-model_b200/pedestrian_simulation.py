"""Per-tick Social Force Model driver with the call surface of the reference's ``pedestrian_simulation.py``.

``tick`` keeps the reference sequence (pedestrian_simulation.py:57-83): mode bookkeeping -> gap acceptance -> recording
-> force sum -> new velocities.  The arithmetic (all enabled forces, their sum in dict order, the Euler velocity update
and the speed clamp) is one fused device pass; as in the reference, positions are left to the simulator that consumes the
velocities (run_simulation.py:77-87).

Three ways through ``tick``, chosen per call:

* **resident** (every ``mode`` entry a stock ``PedModeManager``, the stock force dict, ``record_states=False``): the
  parameters, the point sets, the pedestrian rows and the mode machines stay on the device between ticks; one
  ``sfm_tick_records`` call hands the structured ``PedState.state`` array over as it lies in memory, runs the mode
  bookkeeping (K4a), the forces and the velocity update there and writes the new velocities back into the table.  No
  interpreter loop over pedestrians anywhere; the host objects are refreshed only on ticks on which a machine changed mode.
* **columnar** (same, with ``record_states=True``): the machines tick as array operations on their ``ModeTable``, gap
  acceptance runs on the host for the pedestrians at the kerb, the snapshot is taken where the reference takes it (:76),
  then the same device call without the mode step.
* **generic** (anything else in the ``mode`` column or a hand-composed force dict): the reference's per-object sequence.
"""
import numpy as np

import forces
from check_traffic import check_traffic
from ped_mode_manager import PedMode
from pedestrian_state import PedState
from sfm_b200 import native
from sfm_b200.session import get_session

_NATIVE_ORDER = {name: k for k, name in enumerate(native.FORCE_CLASSES)}


class PedestrianSimulation:
    def __init__(self, borders, border_section_info, obstacles, sfm_config, step_length, record_states=True):
        self.sfm_config = sfm_config
        self.borders = borders
        self.section_info = border_section_info
        self.static_obstacles = obstacles
        self.dyn_obs_ids = []
        self.dyn_obstacles = []
        self.dyn_obs_heading = []
        self.dyn_obs_vel = []
        self.dyn_obs_extent = []
        self.all_dyn_obs_states = {}
        self.record_states = record_states          # False skips the O(N) per-tick snapshots (unbounded memory)
        self.peds = PedState(sfm_config)
        self.step_length = step_length
        self.forces = self.init_forces()
        self.new_velocities = None
        self._dyn_version = 0                       # bumped by update_dynamic_obstacles: the device traffic set is stale
        self._life_counters = None

    def init_forces(self):
        """Ordered dict of the enabled force objects (pedestrian_simulation.py:32-55); order = summation order."""
        switches = self.sfm_config['forces']
        out = {}
        if switches.get('acceleration_force', False):
            out['acceleration_force'] = forces.AccelerationForce(self.step_length, self.sfm_config)
        if switches.get('pedestrian_force', False):
            out['pedestrian_force'] = forces.PedestrianForce(self.step_length, self.sfm_config)
        if switches.get('border_force', False):
            out['border_force'] = forces.BorderForce(self.step_length, self.sfm_config, self.borders, self.section_info)
        if switches.get('static_obstacle_force', False):
            out['static_obstacle_force'] = forces.ObstacleForce(self.step_length, self.sfm_config)
            if self.static_obstacles:
                out['static_obstacle_force'].update_obstacles(self.static_obstacles)
        if switches.get('dynamic_obstacle_force', False):
            out['dynamic_obstacle_force'] = forces.ObstacleForce(self.step_length, self.sfm_config, True)
        for dead in ('ped_repulsive_force', 'space_repulsive_force'):
            if switches.get(dead, False):     # the reference would raise AttributeError here
                raise AttributeError(f"module 'forces' has no class for '{dead}' (dead switch in the reference too)")
        return out

    # ---- the tick ---------------------------------------------------------------------------------------------------
    def tick(self, sim_time):
        """Do one step in the simulation."""
        if self.peds.state is None or self.peds.size() == 0:
            return
        # resident path with the device holding this table's identity column: the object pointers are compared there
        # (sfm_tick_records, identity 'check') instead of by a strided pass over the table on the host
        session = get_session() if not self.record_states else None
        table = self.peds.cached_mode_table() if session is not None and session.identity_table is not None else None
        deferred = table is not None and session.identity_table is table and session.resident_table is table
        if not deferred:
            table = self.peds.mode_table()
        if table is not None and self._fusable() and self._uniform_waiting_time(table):
            if self._tick_table(table, sim_time, check_identity=deferred):
                return
            table = self.peds.mode_table()          # the device saw other objects: full look at the column, then again
            if table is not None and self._fusable() and self._uniform_waiting_time(table):
                self._tick_table(table, sim_time)
                return
        # generic path: apply_current_mode, then every machine's tick (pedestrian_simulation.py:63-65) -- one pass over the
        # mode objects, which also yields the uint8 codes the device needs
        state = self.peds.state
        n = len(state)
        speeds, codes = np.empty(n), np.empty(n, dtype=np.uint8)
        for k, mode in enumerate(state['mode']):
            if hasattr(mode, 'tick'):
                speeds[k] = mode.target_speed
                mode.tick(sim_time)
                codes[k] = mode.current_mode
            else:                                  # plain PedMode / int in the mode column
                speeds[k] = state['target_speed'][k]
                codes[k] = int(mode)
        state['target_speed'] = speeds
        for row in np.nonzero(codes == PedMode.CHECKING_TRAFFIC)[0]:
            ped = state[row]
            ready = True
            if self.dyn_obstacles:
                ready = check_traffic(ped, self.dyn_obstacles, self.dyn_obs_vel, self.dyn_obs_extent)
            if ready:
                ped['mode'].set_mode(PedMode.CROSSING_ROAD)
                codes[row] = ped['mode'].current_mode
        self._mode_codes = codes
        if self.record_states:
            self.peds.record_current_state(sim_time)
            if self.dyn_obstacles:
                self.record_dyn_obstacle_states(sim_time)
        if self._fusable():
            self._fused_velocities()
        else:                                  # hand-composed force dicts: per-class device forces, summed on the host
            self.calculate_new_velocities(sum(f.get_force(self.peds) for f in self.forces.values()))

    @staticmethod
    def _uniform_waiting_time(table):
        w = table.columns['waiting_time']
        return bool((w == w[0]).all())

    def _bind_device(self, session):
        """Parameters and point sets of this simulation resident on the device -- uploads only what is not there yet."""
        params = native.params_from_config(self.sfm_config, self.step_length,
                                           enable={name: name in self.forces for name in native.FORCE_CLASSES})
        session.set_params(params)
        for f in self.forces.values():
            empty_set = (isinstance(f, forces.ObstacleForce) and (f.obstacle_locs is None or f.obstacle_locs.size == 0)) \
                or (isinstance(f, forces.BorderForce) and len(f.borders) == 0)
            if isinstance(f, forces.ObstacleForce) and not empty_set and f.obstacle_velocities is None:
                f.obstacle_velocities = np.zeros((len(f.obstacle_locs), 2))
            if empty_set:
                self._clear_set(session, f.force_class)
            else:
                f._bind(session)

    def _tick_table(self, table, sim_time, check_identity=False):
        """The resident / columnar paths (module docstring): no interpreter loop over pedestrians.  Returns False when
        ``check_identity`` was asked for and the device found other objects in the ``mode`` column than ``table`` was
        built from (nothing has been done then)."""
        state = self.peds.state
        session = get_session()
        ctx = session.ctx
        self._bind_device(session)
        cols = table.columns
        session.pin(state)
        on_device = not self.record_states                  # the mode bookkeeping of :63-73 runs in K4a
        if session.resident_table is not table or ctx.n != len(state):
            ctx.upload_state(*self.peds.device_columns(cols['current_mode']))
            session.resident_table, session.machines_version, session.traffic_version = table, None, None
            session.identity_table = None
            self._life_counters = None
            check_identity = False
        if not on_device:
            # columnar host path: apply_current_mode, tick, gap acceptance for the pedestrians at the kerb, snapshot
            state['target_speed'] = cols['target_speed']                                      # :63
            table.tick(sim_time)                                                               # :64-65
            waiting = np.nonzero(cols['current_mode'] == PedMode.CHECKING_TRAFFIC)[0]
            if len(waiting) and self.dyn_obstacles:                                            # :67-73
                waiting = [r for r in waiting
                           if check_traffic(state[r], self.dyn_obstacles, self.dyn_obs_vel, self.dyn_obs_extent)]
            if len(waiting):
                table.request_crossing(np.asarray(waiting, dtype=np.int64))
            self.peds.record_current_state(sim_time)                                          # :76
            if self.dyn_obstacles:
                self.record_dyn_obstacle_states(sim_time)
            ctx.update_targets(mode=cols['current_mode'])
            ctx.tick_records(state, sim_time, tick_modes=False)
        else:
            if session.machines_version != table.version:   # a fresh table, or writes through the objects (set_mode)
                ctx.update_targets(mode=cols['current_mode'])
                ctx.set_mode_machines(cols['initial_target_speed'], cols['crossing_speed'], cols['crossing_safety_margin'],
                                      cols['target_speed'], cols['next_mode_time'], float(cols['waiting_time'][0]))
                session.machines_version = table.version
            if session.traffic_version != self._dyn_version:
                centres = [c for c, _ in self.dyn_obstacles]
                ctx.set_traffic(centres, self.dyn_obs_vel if len(centres) else [], self.dyn_obs_extent if len(centres) else [])
                session.traffic_version = self._dyn_version
            before = self._life_counters
            if check_identity and session.identity_table is table:
                counters, changed = ctx.tick_records(state, sim_time, tick_modes=True, identity=('check', 'mode'))
                if changed:
                    session.identity_table = None
                    return False
            else:
                counters = ctx.tick_records(state, sim_time, tick_modes=True, identity=('adopt', 'mode'))
                session.identity_table = table
            cols['sim_time'][:] = sim_time
            if before is None or counters[0] != before[0] or counters[1] != before[1]:
                m = ctx.download_modes()                    # somebody started crossing or woke up: refresh the host objects
                cols['current_mode'][:] = m['mode']
                cols['target_speed'][:] = m['mode_target_speed']
                cols['next_mode_time'][:] = m['next_mode_time']
            self._life_counters = counters
        self.new_velocities = state[['id', 'vel']]          # a view: the device call wrote state['vel']
        return True

    def _fusable(self):
        names = list(self.forces)
        return all(n in _NATIVE_ORDER and isinstance(f, forces.Force) and f.force_class == _NATIVE_ORDER[n]
                   for n, f in self.forces.items()) and names == sorted(names, key=_NATIVE_ORDER.get)

    def _fused_velocities(self):
        session = get_session()
        self._bind_device(session)
        session.upload_peds(self.peds, getattr(self, '_mode_codes', None))
        session.ctx.step(1, integrate_positions=False)
        _, vel = session.ctx.download_state()
        self.new_velocities = self.peds.state[['id', 'vel']]     # a view: writes through to state['vel']
        self.new_velocities['vel'] = vel

    @staticmethod
    def _clear_set(session, which):
        if session.owner[which] is not None or session.set_version[which] != 'empty':
            if which == native.BORDER:
                session.ctx.set_borders([], None, None)
            else:
                session.ctx.set_obstacles(which, None, [])
            session.owner[which], session.set_version[which] = None, 'empty'

    def calculate_new_velocities(self, force):
        """New desired velocities from a force array (pedestrian_simulation.py:117-124) -- on the device too
        (``sfm_apply_force``: v + dt F, then the speed clamp of stateutils.py:18-23, in numpy's operation order)."""
        session = get_session()
        session.set_params(native.params_from_config(self.sfm_config, self.step_length,
                                                     enable={name: name in self.forces for name in native.FORCE_CLASSES}))
        session.upload_peds(self.peds, getattr(self, '_mode_codes', None))
        desired_velocity = session.ctx.apply_force(np.ascontiguousarray(force, dtype=np.float64))
        self.new_velocities = self.peds.state[['id', 'vel']]
        self.new_velocities['vel'] = desired_velocity

    def get_new_velocities(self):
        return self.new_velocities

    # ---- bookkeeping (pedestrian_simulation.py:85-115,126-143) ------------------------------------------------------
    def close(self):
        pass

    def get_arrived_peds(self, distance_threshold):
        if self.peds.state is None:
            return []
        to_goal = self.peds.next_waypoint()[:, :2] - self.peds.loc()[:, :2]
        return self.peds.name()[np.linalg.norm(to_goal, axis=-1) < distance_threshold]

    def spawn_pedestrian(self, initial_ped_state):
        self.peds.add_pedestrian(initial_ped_state)

    def destroy_pedestrian(self, ped_name):
        self.peds.remove_pedestrian(ped_name)

    def update_ped_info(self, walker_id, location, velocity):
        self.peds.update_state(walker_id, location, velocity)

    def update_dynamic_obstacles(self, dynamic_obstacles):
        (self.dyn_obs_ids, obstacle_pos, self.dyn_obs_heading, self.dyn_obs_vel, self.dyn_obs_extent,
         borders) = dynamic_obstacles
        self.dyn_obstacles = list(zip(obstacle_pos, borders))
        self._dyn_version += 1
        if 'dynamic_obstacle_force' in self.forces and self.dyn_obstacles:
            self.forces['dynamic_obstacle_force'].update_obstacles(self.dyn_obstacles)
            self.forces['dynamic_obstacle_force'].update_obstacle_velocities(self.dyn_obs_vel)

    def record_dyn_obstacle_states(self, sim_time):
        centres = [c for c, _ in self.dyn_obstacles]
        veh_state = np.empty(len(self.dyn_obs_ids), dtype=[('id', 'i4'), ('loc', 'f8', (2,)), ('heading', 'f8'),
                                                           ('vel', 'f8', (2,)), ('extent', 'f8', (2,))])
        veh_state['id'], veh_state['loc'], veh_state['heading'] = self.dyn_obs_ids, centres, self.dyn_obs_heading
        veh_state['vel'], veh_state['extent'] = self.dyn_obs_vel, self.dyn_obs_extent
        self.all_dyn_obs_states[sim_time] = veh_state

    def get_states(self):
        return self.peds.get_all_states()
