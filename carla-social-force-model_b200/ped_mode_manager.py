"""Pedestrian modes and the per-pedestrian mode state machine (call surface of the reference's ``ped_mode_manager.py``).

The device path only consumes the two quantities a mode determines -- ``target_speed`` and whether the border force is
masked (modes CROSSING_ROAD / ROAD_TO_SIDEWALK, forces.py:176-177) -- as a ``uint8`` and a ``float64`` column; the same
machine runs on the device as K4a (``csrc/k4_lifecycle.cuh``).  Here it is table driven: what a mode does to the target
speed is data, and so are the two detours a request can take.

A ``PedModeManager`` is an ordinary Python object, exactly as the reference's callers expect
(pedestrian_spawner.py:238-241 builds one per pedestrian, run_simulation.py:123-125 calls ``set_mode`` on it).  Its
fields, however, can live in a ``ModeTable`` -- one numpy column per field for all the managers of a pedestrian table --
so that ``PedestrianSimulation.tick`` advances a crowd of machines with array operations (or on the device) instead of
one interpreter iteration per pedestrian (pedestrian_simulation.py:63-65).  Adoption is transparent: the object keeps
answering attribute reads and writes, now backed by its row of the table.
"""
from enum import IntEnum

import numpy as np


class PedMode(IntEnum):                      # ped_mode_manager.py:4-9
    IDLE = 0
    WALKING_SIDEWALK = 1
    CROSSING_ROAD = 2
    ROAD_TO_SIDEWALK = 3
    CHECKING_TRAFFIC = 4


# requested transition (from, to) -> the intermediate mode activated instead (ped_mode_manager.py:42-47)
_DETOURS = {
    (PedMode.WALKING_SIDEWALK, PedMode.CROSSING_ROAD): PedMode.CHECKING_TRAFFIC,
    (PedMode.CROSSING_ROAD, PedMode.WALKING_SIDEWALK): PedMode.ROAD_TO_SIDEWALK,
}

# mode -> attribute holding the target speed the mode imposes; 0 = stand still, None = keep the current speed (:49-69)
_SPEED_OF = {
    PedMode.IDLE: 0,
    PedMode.WALKING_SIDEWALK: 'initial_target_speed',
    PedMode.CROSSING_ROAD: 'crossing_speed',
    PedMode.ROAD_TO_SIDEWALK: None,
    PedMode.CHECKING_TRAFFIC: 0,
}

_FLOAT_FIELDS = ('sim_time', 'next_mode_time', 'initial_target_speed', 'target_speed', 'crossing_speed',
                 'crossing_safety_margin', 'waiting_time')


def _column_property(name):
    def fget(self):
        table = self._table
        if table is None:
            return self._own[name]
        return table.columns[name][self._row].item()

    def fset(self, value):
        table = self._table
        if table is None:
            self._own[name] = value
        else:
            table.columns[name][self._row] = value
            table.version += 1                     # host-side write: device mirrors of the table are stale
    return property(fget, fset)


class PedModeManager:
    """Finite state machine deciding a pedestrian's mode and the target speed that goes with it."""

    __slots__ = ('ped_name', '_own', '_table', '_row')

    def __init__(self, ped_name, target_speed, initial_mode, crossing_speed_factor, crossing_safety_margin):
        self.ped_name = ped_name
        self._table, self._row = None, -1
        self._own = dict(sim_time=0, next_mode_time=-1, initial_target_speed=target_speed, target_speed=target_speed,
                         crossing_speed=crossing_speed_factor * target_speed,
                         crossing_safety_margin=crossing_safety_margin, waiting_time=5, current_mode=initial_mode)

    sim_time = _column_property('sim_time')
    next_mode_time = _column_property('next_mode_time')
    initial_target_speed = _column_property('initial_target_speed')
    target_speed = _column_property('target_speed')
    crossing_speed = _column_property('crossing_speed')
    crossing_safety_margin = _column_property('crossing_safety_margin')
    waiting_time = _column_property('waiting_time')       # seconds an IDLE pedestrian waits before it walks (:28)

    @property
    def current_mode(self):
        table = self._table
        if table is None:
            return self._own['current_mode']
        return PedMode(int(table.columns['current_mode'][self._row]))

    @current_mode.setter
    def current_mode(self, mode):
        table = self._table
        if table is None:
            self._own['current_mode'] = mode
        else:
            table.columns['current_mode'][self._row] = int(mode)
            table.version += 1

    def tick(self, sim_time):
        """Advance to ``sim_time``; an idle pedestrian starts walking once its waiting time is over (:30-35)."""
        self.sim_time = sim_time
        if self.current_mode == PedMode.IDLE and self.next_mode_time <= sim_time:
            self._activate_mode(PedMode.WALKING_SIDEWALK)

    def set_mode(self, new_mode):
        """Request ``new_mode``; sidewalk->road goes through CHECKING_TRAFFIC, road->sidewalk through ROAD_TO_SIDEWALK."""
        self._activate_mode(_DETOURS.get((self.current_mode, new_mode), new_mode))

    def _activate_mode(self, mode):
        if mode not in _SPEED_OF:
            return                                 # unknown mode: ignored, like the reference's if/elif chain
        rule = _SPEED_OF[mode]
        if rule is not None:
            self.target_speed = getattr(self, rule) if isinstance(rule, str) else rule
        if mode == PedMode.IDLE:
            self.next_mode_time = self.sim_time + self.waiting_time
        self.current_mode = mode


class ModeTable:
    """Columnar storage of the mode machines of one pedestrian table (row k = pedestrian k).

    ``columns[name]`` is a float64 array per PedModeManager field plus the uint8 ``current_mode``; ``version`` counts
    writes that came through the objects (``set_mode`` from the waypoint hand-over, run_simulation.py:123-125), so a
    device mirror knows when to refresh.  ``tick`` is PedModeManager.tick for every row at once.
    """

    def __init__(self, managers):
        n = len(managers)
        self.managers = managers
        self.columns = {name: np.empty(n, dtype=np.float64) for name in _FLOAT_FIELDS}
        self.columns['current_mode'] = np.empty(n, dtype=np.uint8)
        cols = self.columns
        for k, m in enumerate(managers):
            if m._table is not None:               # still bound to an older table (e.g. after a despawn): take it out
                m._table.detach(m)
            own = m._own
            for name in _FLOAT_FIELDS:
                cols[name][k] = own[name]
            cols['current_mode'][k] = int(own['current_mode'])
            m._table, m._row = self, k
        self.version = 0

    @staticmethod
    def adoptable(modes):
        return all(type(m) is PedModeManager for m in modes)

    def detach(self, m):
        """Give manager ``m`` its values back (it leaves the pedestrian table or the table is being rebuilt)."""
        k = m._row
        own = {name: self.columns[name][k].item() for name in _FLOAT_FIELDS}
        own['current_mode'] = PedMode(int(self.columns['current_mode'][k]))
        m._own, m._table, m._row = own, None, -1

    def release(self):
        for m in self.managers:
            if m._table is self:
                self.detach(m)

    def tick(self, sim_time):
        """ped_mode_manager.py:30-35 for every machine: returns the number of idle pedestrians that woke up."""
        c = self.columns
        c['sim_time'][:] = sim_time
        wake = (c['current_mode'] == PedMode.IDLE) & (c['next_mode_time'] <= sim_time)
        if wake.any():
            c['target_speed'][wake] = c['initial_target_speed'][wake]
            c['current_mode'][wake] = PedMode.WALKING_SIDEWALK
        return int(wake.sum())

    def request_crossing(self, rows):
        """``set_mode(CROSSING_ROAD)`` for pedestrians waiting at the kerb (pedestrian_simulation.py:72-73)."""
        c = self.columns
        c['target_speed'][rows] = c['crossing_speed'][rows]
        c['current_mode'][rows] = PedMode.CROSSING_ROAD
