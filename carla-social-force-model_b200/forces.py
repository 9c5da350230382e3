"""Social forces with the call surface of the reference's ``forces.py``, evaluated by the sm_100a kernels.

Every class keeps the reference constructor and ``get_force(ped_state, debug=False) -> (N, 3) float64`` contract
(forces.py:11-32) and reads the same ``sfm_config`` keys with the same defaults and the same mandatory sections
(``KeyError`` where the reference raises one).  ``_get_force`` uploads the pedestrian table and asks the device context
for that class's force (``sfm_force`` in include/sfm_b200.h); there is no CPU implementation behind these classes.
"""
import logging
from abc import ABC, abstractmethod

import numpy as np

from sfm_b200 import native
from sfm_b200.session import get_session


class Force(ABC):
    """Force base class (forces.py:11-32)."""

    def __init__(self, step_length, sfm_config):
        super().__init__()
        self.step_length = step_length
        self.sfm_config = sfm_config
        self.use_ped_radius = sfm_config.get('use_ped_radius', False)

    @abstractmethod
    def _get_force(self, ped_state):
        raise NotImplementedError

    def get_force(self, ped_state, debug=False):
        force = self._get_force(ped_state)
        if debug:
            logging.debug(f"{type(self).__name__}:\n {repr(force)}")
        return force

    # -- device plumbing shared by the concrete classes
    force_class = None

    def _bind(self, session):
        """Hook for classes that own a point set."""

    def _device_force(self, peds):
        if peds.size() == 0:
            return np.zeros((0, 3))
        session = get_session()
        session.bind_params(self.sfm_config, self.step_length)
        self._bind(session)
        session.upload_peds(peds)
        return session.ctx.force(self.force_class)


class AccelerationForce(Force):
    """Relaxation towards the desired velocity, Helbing & Molnar 1995 (forces.py:35-53)."""
    force_class = native.ACCELERATION

    def __init__(self, step_length, sfm_config):
        super().__init__(step_length, sfm_config)
        self.tau = self.sfm_config.get('goal_force', {}).get('tau', 0.5)          # sic: [goal_force], forces.py:44

    def _get_force(self, peds):
        return self._device_force(peds)


class _MoussaidMixin:
    def _read_moussaid(self, section):
        self.lambda_weight = section.get('lambda', 2.0)
        self.A = section.get('A', 4.5)
        self.gamma = section.get('gamma', 0.35)
        self.n = section.get('n', 2.0)
        self.n_prime = section.get('n_prime', 3.0)
        self.epsilon = section.get('epsilon', 0.005)


class PedestrianForce(Force, _MoussaidMixin):
    """Pedestrian-pedestrian interaction, Moussaid et al. 2009 (forces.py:56-117) -- the all-pairs kernel K1."""
    force_class = native.PEDESTRIAN

    def __init__(self, step_length, sfm_config):
        super().__init__(step_length, sfm_config)
        self.ped_force_config = self.sfm_config['pedestrian_force']              # mandatory section, forces.py:66
        self._read_moussaid(self.ped_force_config)

    def _get_force(self, peds):
        return self._device_force(peds)


class BorderForce(Force):
    """Repulsion from the nearest point of every close border section (forces.py:120-179) -- kernels K2a/K2b."""
    force_class = native.BORDER

    def __init__(self, step_length, sfm_config, borders, section_info):
        super().__init__(step_length, sfm_config)
        self.borders = borders
        if len(borders):
            info = np.asarray(section_info, dtype=object)                          # rows [centre(2), length]
            self.section_center = np.vstack(list(info[:, 0])).astype(np.float64)
            self.section_length = np.asarray(info[:, 1], dtype=np.float64)
        else:
            self.section_center, self.section_length = np.zeros((0, 2)), np.zeros(0)
        self.border_force_config = self.sfm_config['border_force']               # mandatory section, forces.py:134
        self.a = self.border_force_config.get('a', 3.0)
        self.b = self.border_force_config.get('b', 0.1)

    def _bind(self, session):
        session.bind_set(native.BORDER, self, len(self.borders),
                         lambda ctx: ctx.set_borders(self.borders, self.section_center, self.section_length))

    def _get_force(self, peds):
        if len(self.borders) == 0:                                                # forces.py:140-141
            return np.zeros((peds.size(), 3))
        return self._device_force(peds)


class ObstacleForce(Force, _MoussaidMixin):
    """Moussaid interaction with the nearest ring point of every close obstacle (forces.py:182-291) -- kernels K2a/K2c."""

    def __init__(self, step_length, sfm_config, dynamic=False):
        super().__init__(step_length, sfm_config)
        self.obstacle_locs = None
        self.obstacle_borders = None
        self.obstacle_velocities = None
        self.dynamic = dynamic
        self.force_class = native.DYNAMIC_OBSTACLE if dynamic else native.STATIC_OBSTACLE
        key = 'dynamic_obstacle_force' if dynamic else 'static_obstacle_force'  # mandatory, forces.py:197,199
        self.evasion_force_config = self.sfm_config[key]
        self._read_moussaid(self.evasion_force_config)
        self.perception_threshold = self.evasion_force_config.get('perception_threshold', 20)
        self._version = 0

    def update_obstacles(self, obstacles):
        """``obstacles`` = iterable of (centre(2), ring(P, 2)) -- forces.py:285-288."""
        obstacle_locs, obstacle_borders = zip(*obstacles)
        self.obstacle_locs = np.array(obstacle_locs)
        self.obstacle_borders = obstacle_borders
        self._version += 1

    def update_obstacle_velocities(self, velocities):
        self.obstacle_velocities = np.array(velocities)                           # forces.py:290-291
        self._version += 1

    def _bind(self, session):
        def load(ctx):
            ctx.set_obstacles(self.force_class, self.obstacle_locs, self.obstacle_borders, self.obstacle_velocities)
        session.bind_set(self.force_class, self, self._version, load)

    def _get_force(self, peds):
        if self.obstacle_locs is None or self.obstacle_locs.size == 0:            # forces.py:209-210
            return np.zeros((peds.size(), 3))
        if self.obstacle_velocities is None:                                      # forces.py:212-213
            self.obstacle_velocities = np.zeros((len(self.obstacle_locs), 2))
        return self._device_force(peds)
