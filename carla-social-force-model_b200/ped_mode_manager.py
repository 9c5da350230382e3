"""Pedestrian modes and the per-pedestrian mode state machine (call surface of the reference's ``ped_mode_manager.py``).

The device path only consumes the two quantities a mode determines -- ``target_speed`` and whether the border force is
masked (modes CROSSING_ROAD / ROAD_TO_SIDEWALK, forces.py:176-177) -- as a ``uint8`` and a ``float64`` column; the same
machine runs on the device as K4a (``csrc/k4_lifecycle.cuh``) for device-resident crowds.  Here it is table driven:
what a mode does to the target speed is data, and so are the two detours a request can take.
"""
from enum import IntEnum


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


class PedModeManager:
    """Finite state machine deciding a pedestrian's mode and the target speed that goes with it."""

    waiting_time = 5                               # seconds an IDLE pedestrian waits before it starts walking (:28)

    def __init__(self, ped_name, target_speed, initial_mode, crossing_speed_factor, crossing_safety_margin):
        self.ped_name = ped_name
        self.sim_time = 0
        self.next_mode_time = -1
        self.initial_target_speed = self.target_speed = target_speed
        self.crossing_speed = crossing_speed_factor * target_speed
        self.crossing_safety_margin = crossing_safety_margin
        self.current_mode = initial_mode
        self.waiting_time = PedModeManager.waiting_time

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
