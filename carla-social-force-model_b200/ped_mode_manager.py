"""Pedestrian modes and the per-pedestrian mode state machine (call surface of the reference's ``ped_mode_manager.py``).

The device path only consumes the two quantities a mode determines -- ``target_speed`` and whether the border force is
masked (modes CROSSING_ROAD / ROAD_TO_SIDEWALK, forces.py:176-177) -- as a ``uint8`` and a ``float64`` column.
"""
from enum import IntEnum


class PedMode(IntEnum):                      # ped_mode_manager.py:4-9
    IDLE = 0
    WALKING_SIDEWALK = 1
    CROSSING_ROAD = 2
    ROAD_TO_SIDEWALK = 3
    CHECKING_TRAFFIC = 4


# intermediate mode inserted when a transition is requested: (current, requested) -> activated  (ped_mode_manager.py:42-47)
_DETOURS = {
    (PedMode.WALKING_SIDEWALK, PedMode.CROSSING_ROAD): PedMode.CHECKING_TRAFFIC,
    (PedMode.CROSSING_ROAD, PedMode.WALKING_SIDEWALK): PedMode.ROAD_TO_SIDEWALK,
}


class PedModeManager:
    """Finite state machine deciding a pedestrian's mode and the target speed that goes with it."""

    def __init__(self, ped_name, target_speed, initial_mode, crossing_speed_factor, crossing_safety_margin):
        self.ped_name = ped_name
        self.sim_time = 0
        self.initial_target_speed = target_speed
        self.crossing_speed = crossing_speed_factor * target_speed
        self.crossing_safety_margin = crossing_safety_margin
        self.waiting_time = 5
        self.next_mode_time = -1
        self.current_mode = initial_mode
        self.target_speed = target_speed

    def tick(self, sim_time):
        """Advance to ``sim_time``; an idle pedestrian starts walking once its waiting time is over (:30-35)."""
        self.sim_time = sim_time
        if self.current_mode == PedMode.IDLE and self.next_mode_time <= sim_time:
            self._activate_mode(PedMode.WALKING_SIDEWALK)

    def set_mode(self, new_mode):
        """Request ``new_mode``; sidewalk->road goes through CHECKING_TRAFFIC, road->sidewalk through ROAD_TO_SIDEWALK."""
        self._activate_mode(_DETOURS.get((self.current_mode, new_mode), new_mode))

    def _activate_mode(self, mode):
        # target speed per mode (:49-69); ROAD_TO_SIDEWALK keeps whatever speed was active
        if mode == PedMode.IDLE:
            self.target_speed = 0
            self.next_mode_time = self.sim_time + self.waiting_time
        elif mode == PedMode.WALKING_SIDEWALK:
            self.target_speed = self.initial_target_speed
        elif mode == PedMode.CROSSING_ROAD:
            self.target_speed = self.crossing_speed
        elif mode == PedMode.CHECKING_TRAFFIC:
            self.target_speed = 0
        elif mode != PedMode.ROAD_TO_SIDEWALK:
            return                                        # unknown mode: ignored, like the reference's if/elif chain
        self.current_mode = mode
