"""CPU timing of the oracle on a bounded row sample -- the reported CPU baseline of bench.py.  TEST INFRASTRUCTURE.

The reference is single-threaded numpy and cannot run N >= ~8k at all (dense (N, N-1, 3) temporaries); the only CPU
number available at the benchmark sizes is the row-chunked float64 restatement (``kind = "port"``), run here on every
host core with one forked worker per core, each computing the full tick (all enabled force classes + velocity update)
for its share of a row sample.  Throughput is extrapolated by rows: agent-steps/s = sample rows / wall time.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np

from . import sfm_oracle as O

_SHARED = {}
_WORKER_BYTES = 0.7e9          # budget of numpy temporaries per worker (the pair force builds ~20 row-chunk x N arrays)


def _chunk_rows(n):
    """Rows per pedestrian-force chunk so that one worker's temporaries stay within the budget (~290 B per pair)."""
    return int(min(48, max(1, _WORKER_BYTES * 0.85 / (n * 290.0))))


def _usable_cores(cores):
    """All host cores, unless a quarter of the box's memory cannot hold one worker's temporaries per core."""
    try:
        with open('/proc/meminfo') as f:
            total = next(int(line.split()[1]) * 1024 for line in f if line.startswith('MemTotal'))
        return max(1, min(cores, int(0.25 * total / _WORKER_BYTES)))
    except Exception:
        return min(cores, 16)


def _worker(rows):
    w, scene, dyn, dyn_vel = _SHARED['w'], _SHARED['scene'], _SHARED['dyn'], _SHARED['dyn_vel']
    t0 = time.perf_counter()
    per_class = O.forces_by_class(scene, w.loc, w.vel, w.next_waypoint, w.radius, w.target_speed, w.mode, dyn, dyn_vel,
                                  rows=rows, chunk=_chunk_rows(w.n))
    F = O.total_force(per_class, len(rows))
    O.new_velocities(w.vel[rows], F, w.target_speed[rows], scene.dt, scene.max_speed_factor)
    return time.perf_counter() - t0


def time_sample(w, sfm_config, rows_per_core=64, cores=None, repeats=1, seed=0):
    """Returns dict(seconds, rows, cores, agent_steps_per_s, pairs_per_s) for one tick over a random row sample."""
    cores = _usable_cores(cores or os.cpu_count() or 1)
    scene = O.Scene(sfm_config, w.step_length, w.borders, w.section_center, w.section_length, w.static_obstacles)
    veh = w.vehicles_at(0)
    dyn, dyn_vel = (list(zip(veh[1], veh[5])), veh[3]) if veh is not None else (None, None)
    _SHARED.update(w=w, scene=scene, dyn=dyn, dyn_vel=dyn_vel)
    rng = np.random.default_rng(seed)
    n_rows = min(w.n, rows_per_core * cores)
    rows = np.sort(rng.choice(w.n, size=n_rows, replace=False))
    chunks = [c for c in np.array_split(rows, cores) if len(c)]
    best = None
    if cores == 1:
        for _ in range(repeats):
            t0 = time.perf_counter()
            _worker(chunks[0])
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    else:
        ctx = mp.get_context('fork')                 # workers inherit the workload copy-on-write
        with ctx.Pool(len(chunks)) as pool:
            pool.map(_worker, [c[:1] for c in chunks])           # spin the workers up outside the timed region
            for _ in range(repeats):
                t0 = time.perf_counter()
                pool.map(_worker, chunks)
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
    return dict(seconds=best, rows=int(n_rows), cores=len(chunks), agent_steps_per_s=n_rows / best,
                pairs_per_s=n_rows * (w.n - 1) / best)
