"""Pedestrian state table with the call surface of the reference's ``pedestrian_state.py``.

``PedState.state`` is the same structured numpy array the reference keeps (dtype pedestrian_state.py:17-19) because
its callers index it directly (pedestrian_simulation.py:64,67,123; check_traffic.py:18-21; run_simulation.py via
``update_state``).  It is the *host mirror*: the device keeps its own SoA copy (float64 master + float32 staging, see
DESIGN.md) that ``device_columns()`` feeds.
"""
import ctypes

import numpy as np

import stateutils
from ped_mode_manager import ModeTable, PedMode


class PedState:
    def __init__(self, sfm_config):
        self.max_speed_factor = sfm_config.get('max_speed_factor', 1.3)      # sic: not `max_speed_multiplier`
        self.ped_state_dtype = [('name', 'U8'), ('id', 'i4'), ('loc', 'f8', (3,)), ('vel', 'f8', (3,)),
                                ('next_waypoint', 'f8', (3,)), ('mode', 'O'), ('radius', 'f8'), ('target_speed', 'f8')]
        self.state = None
        self.all_states = {}            # sim_time -> snapshot, filled by record_current_state
        self._table, self._table_ptrs = None, None      # ModeTable of the current rows + the objects it was built from
        self._table_state = None                        # ... and the state array that comparison was made on

    # ---- rows in / out (pedestrian_state.py:26-43) ------------------------------------------------------------------
    def add_pedestrian(self, initial_ped_state):
        row = np.array([tuple(initial_ped_state)], dtype=self.ped_state_dtype)
        self.state = row if self.state is None else np.concatenate((self.state, row))

    def add_pedestrians(self, names, ids, loc, vel, next_waypoint, modes, radius, target_speed):
        """Bulk insert (not in the reference): one allocation instead of one ``np.append`` per pedestrian."""
        block = np.zeros(len(ids), dtype=self.ped_state_dtype)
        block['name'], block['id'] = names, ids
        block['loc'], block['vel'], block['next_waypoint'] = loc, vel, next_waypoint
        block['radius'], block['target_speed'] = radius, target_speed
        block['mode'] = list(modes)
        self.state = block if self.state is None else np.concatenate((self.state, block))

    def remove_pedestrian(self, ped_name):
        self.state = self.state[self.state['name'] != ped_name]

    # ---- column accessors (pedestrian_state.py:45-77) ---------------------------------------------------------------
    def size(self):
        return self.state.shape[0]

    def name(self):
        return self.state['name']

    def walker_id(self):
        return self.state['id']

    def loc(self):
        return self.state['loc']

    def vel(self):
        return self.state['vel']

    def next_waypoint(self):
        return self.state['next_waypoint']

    def mode(self):
        return self.state['mode']

    def radius(self):
        return self.state['radius']

    def target_speed(self):
        return self.state['target_speed']

    def max_speed(self):
        return self.target_speed() * self.max_speed_factor

    def speeds(self):
        return stateutils.speeds(self.state)

    def desired_directions(self):
        return stateutils.desired_directions(self.state)

    # ---- per-tick updates (pedestrian_state.py:79-95) ---------------------------------------------------------------
    def update_state(self, walker_id, location, velocity):
        rows = self.state['id'] == walker_id
        self.state['loc'][rows] = location
        self.state['vel'][rows] = velocity

    def update_states(self, walker_ids, locations, velocities):
        """Batched ``update_state`` (not in the reference): one id->row lookup instead of an O(N) mask per walker."""
        order = np.argsort(self.state['id'], kind='stable')
        rows = order[np.searchsorted(self.state['id'], walker_ids, sorter=order)]
        self.state['loc'][rows] = locations
        self.state['vel'][rows] = velocities

    def update_next_waypoint(self, ped_name, next_waypoint_tuple):
        next_waypoint, crossing_road = next_waypoint_tuple
        rows = self.state['name'] == ped_name
        self.state['next_waypoint'][rows] = next_waypoint
        wanted = PedMode.CROSSING_ROAD if crossing_road else PedMode.WALKING_SIDEWALK
        self.state['mode'][rows][0].set_mode(wanted)

    def apply_current_mode(self):
        table = self.mode_table()
        if table is not None:
            self.state['target_speed'] = table.columns['target_speed']
            return
        self.state['target_speed'] = [getattr(m, 'target_speed', s)
                                      for m, s in zip(self.state['mode'], self.state['target_speed'])]

    def mode_codes(self):
        """``uint8`` mode per pedestrian; the ``mode`` column may hold ``PedModeManager`` objects or plain ``PedMode`` ints."""
        table = self.mode_table()
        if table is not None:
            return table.columns['current_mode'].copy()
        return np.fromiter((int(getattr(m, 'current_mode', m)) for m in self.state['mode']), dtype=np.uint8,
                           count=self.size())

    # ---- columnar mode machines ---------------------------------------------------------------------------------------
    def _mode_pointers(self):
        """The ``mode`` column as raw object addresses (int32 [n, 2] halves of the 8-byte pointers; the 132-byte records
        keep 4-byte alignment only): identity of every entry without touching the objects, so that one vectorised
        comparison tells whether the column still holds the objects a ModeTable was built from."""
        s = self.state
        n = len(s)
        if n == 0:
            return np.zeros((0, 2), dtype=np.int32)
        offset = s.dtype.fields['mode'][1]
        span = (n - 1) * s.strides[0] + s.dtype.itemsize
        raw = (ctypes.c_char * span).from_address(s.ctypes.data)
        return np.ndarray(shape=(n, 2), dtype=np.int32, buffer=raw, offset=offset, strides=(s.strides[0], 4))

    def mode_table(self):
        """The ``ModeTable`` backing the ``mode`` column, (re)built when the column holds other objects than last time
        (spawn, despawn, a caller replacing ``state``); ``None`` unless every entry is a stock ``PedModeManager``."""
        if self.state is None or len(self.state) == 0 or self.state.strides[0] <= 0:
            return None
        if self._table_ptrs is not None and len(self.state) == len(self._table_ptrs) and self._same_objects():
            self._table_state = self.state
            return self._table
        if self._table is not None:
            self._table.release()
        modes = list(self.state['mode'])
        self._table = ModeTable(modes) if ModeTable.adoptable(modes) else None
        self._table_ptrs = self._mode_pointers().copy()
        self._table_state = self.state
        return self._table

    def cached_mode_table(self):
        """The last ``mode_table()`` result if ``state`` is still the very array it was validated for, else ``None`` -- no
        look at the column: the caller (the resident tick) has the device compare the object pointers instead."""
        if self._table is not None and self.state is not None and self._table_state is self.state:
            return self._table
        return None

    def _same_objects(self):
        try:                                    # strided memcmp in the native library (host code, ~0.05 ms at N = 65,536)
            from sfm_b200 import native
            return native.column_equal(self.state, 'mode', self._table_ptrs.view(np.uint8).reshape(len(self._table_ptrs), 8))
        except Exception:                       # library not built: the same comparison in numpy
            return np.array_equal(self._mode_pointers(), self._table_ptrs)

    def device_columns(self, mode_codes=None):
        """Contiguous float64 / uint8 columns in the order ``sfm_upload_state`` takes them."""
        s = self.state
        codes = self.mode_codes() if mode_codes is None or len(mode_codes) != len(s) else mode_codes
        return (np.ascontiguousarray(s['loc']), np.ascontiguousarray(s['vel']), np.ascontiguousarray(s['next_waypoint']),
                np.ascontiguousarray(s['radius']), np.ascontiguousarray(s['target_speed']), codes)

    # ---- recording (pedestrian_state.py:100-107) --------------------------------------------------------------------
    def record_current_state(self, sim_time):
        snapshot = self.state.copy()
        table = self.mode_table()
        if table is not None:
            snapshot['mode'] = table.columns['current_mode'].astype(object)
        else:
            snapshot['mode'] = [getattr(m, 'current_mode', m) for m in snapshot['mode']]
        self.all_states[sim_time] = snapshot

    def get_all_states(self):
        return self.all_states
