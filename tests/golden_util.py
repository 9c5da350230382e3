"""Helpers shared by the golden-vector tests (load, digest check, oracle scene construction)."""
import os

import numpy as np

from oracle import sfm_oracle as O
from oracle.make_golden import workload_digest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(name, workload):
    g = np.load(os.path.join(GOLDEN, name))
    assert str(g['digest']) == workload_digest(workload), 'synthetic generator drifted; regenerate tests/golden'
    return g


def scene_for(w, cfg):
    return O.Scene(cfg, w.step_length, w.borders, w.section_center, w.section_length, w.static_obstacles)


def dyn_for(w, step):
    veh = w.vehicles_at(step)
    if veh is None:
        return None, None
    return list(zip(veh[1], veh[5])), veh[3]
