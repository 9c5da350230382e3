"""Device-resident crowd engine: one process per GPU, pedestrians row-partitioned across the ranks of one box.

Rows shard: each rank owns the state, the cell-list forces and the integration of a contiguous row block.  The pair
force is evaluated once per unordered pair by exactly one rank (half-shell over 256-row tiles), so a step has two
exchanges: an integer reduce-scatter of the fixed-point force accumulators (32 B per pedestrian) and an all-gather of
each rank's staged block (60 B per pedestrian: 15 float32 planes).  Two transports:

* ``exchange='peer'`` (default on one box): the library maps every rank's buffers through CUDA IPC once, and the two
  collectives are folded into its own kernels -- ``k1_sym_finish`` pulls the partial accumulators over NVLink, K3 pushes
  the staged rows into every peer's gather buffer, flag barriers in between (``csrc/k7_peer.cuh``).  ``torch.distributed``
  only carries the 192-byte handle table at set-up.
* ``exchange='nccl'``: ``reduce_scatter_tensor`` / ``all_gather_into_tensor`` on tensor views of the library's buffers.

torch is plumbing only -- streams, the process group, tensor views; all arithmetic runs in ``libsfm_b200.so``.
"""
from __future__ import annotations

import numpy as np

from . import native

ROW_ALIGN = 256


def partition_rows(n, world):
    """Contiguous row blocks made of whole 256-row tiles of the pair kernel (block sizes differ by at most one tile; the
    last block takes the ragged tail): returns int64 [world + 1] boundaries.  With whole tiles per rank the tile pairs of
    the multi-rank run are those of the single-GPU run, so -- when the tile count divides by the rank count -- the two
    are bit-identical for ANY crowd size (integer accumulation is order-free).  Crowds smaller than one tile per rank are
    split evenly by rows."""
    n, world = int(n), int(world)
    tiles = (n + ROW_ALIGN - 1) // ROW_ALIGN
    if tiles >= world:
        bounds = np.minimum(n, ROW_ALIGN * ((tiles * np.arange(world + 1, dtype=np.int64)) // world))
        bounds[-1] = n
        return bounds.astype(np.int64)
    base, extra = divmod(n, world)
    sizes = np.full(world, base, dtype=np.int64)
    sizes[:extra] += 1
    bounds = np.zeros(world + 1, dtype=np.int64)
    np.cumsum(sizes, out=bounds[1:])
    return bounds


def padded_rows(bounds):
    """Staged rows per rank block: the largest block rounded up to the pair kernel's tile (256 rows)."""
    largest = int(np.max(np.diff(bounds))) if len(bounds) > 1 else 0
    return max(ROW_ALIGN, (largest + ROW_ALIGN - 1) // ROW_ALIGN * ROW_ALIGN)


def crowd_origin(loc):
    """Origin of the float32 staging copies: the centre of the crowd's bounding box rounded to whole metres (so that
    coordinates which are exact in float32 mostly stay exact), z of the first pedestrian (a flat crowd stages z = 0 and
    takes the planar fast path).  Halves the largest staged magnitude -- and with it the float32 rounding of positions --
    compared with an origin at (0, 0)."""
    if len(loc) == 0:
        return 0.0, 0.0, 0.0
    lo, hi = loc[:, :2].min(axis=0), loc[:, :2].max(axis=0)
    cx, cy = np.round((lo + hi) * 0.5)
    return float(cx), float(cy), float(loc[0, 2])


class _DeviceView:
    """Minimal ``__cuda_array_interface__`` holder so torch can alias memory owned by the library."""

    def __init__(self, ptr, n_items, typestr='<f4'):
        self.__cuda_array_interface__ = {'shape': (n_items,), 'typestr': typestr, 'data': (ptr, False), 'version': 2}


class Engine:
    def __init__(self, sfm_config, step_length, device=None, group=None, use_torch_stream=True, exchange=None,
                 reorder_every=32):
        """``reorder_every``: ticks between rebuilds of the staged slot order (rows staged along a Hilbert curve so that the
        pair kernel's one-subtraction local path applies, ``csrc/k8_order.cuh``); 0 keeps rows staged in row order -- then a
        multi-rank run and a single-GPU run of the same crowd agree bit for bit without exchanging the order."""
        import os
        import torch                                   # plumbing only
        self.exchange_mode = exchange or os.environ.get('SFM_EXCHANGE', 'peer')
        if self.exchange_mode not in ('peer', 'nccl'):
            raise ValueError("exchange must be 'peer' or 'nccl'")
        self.torch = torch
        if not torch.cuda.is_available():
            raise native.SfmError('no CUDA device: the Social Force Model step has no CPU fallback')
        self.dist = torch.distributed if (torch.distributed.is_available() and torch.distributed.is_initialized()) \
            else None
        self.group = group
        self.world = self.dist.get_world_size(group) if self.dist else 1
        self.rank = self.dist.get_rank(group) if self.dist else 0
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.ctx = native.Context(self.device)
        self.sfm_config, self.step_length = sfm_config, step_length
        self.ctx.set_params(native.params_from_config(sfm_config, step_length))
        self.reorder_every = int(os.environ.get('SFM_REORDER_EVERY', reorder_every))
        self.ctx.set_reorder_interval(self.reorder_every)
        # A dedicated (non-default) torch stream carries both the library's kernels and the NCCL all-gather, so the
        # two are ordered without host synchronisation; CUDA events recorded on it time the whole step.
        prio = -1 if os.environ.get('SFM_AUX_PRIORITY', 'high') == 'low' else 0      # experiment knob, see sfm_api.cu
        self.stream = torch.cuda.Stream(self.device, priority=prio) if use_torch_stream else None
        if self.stream is not None:
            self.ctx.set_stream(self.stream.cuda_stream)
        self.n_global = 0
        self.bounds = None
        self._gather = None
        self.device_vehicles = False
        self._ticks = 0                     # ticks stepped since load(): sim_time and the vehicle clock

    # ---- set-up ---------------------------------------------------------------------------------------------------
    def load(self, w, device_vehicles=False):
        """Take a ``synth.Workload`` (or anything with the same attributes): this rank keeps its row block."""
        self.n_global = w.n
        self.bounds = partition_rows(w.n, self.world)
        lo, hi = int(self.bounds[self.rank]), int(self.bounds[self.rank + 1])
        if self.world > 1:
            self.ctx.set_partition(self.world, self.rank, padded_rows(self.bounds))
        self.ctx.set_origin(*crowd_origin(w.loc))          # one origin for every rank
        self.ctx.upload_state(w.loc[lo:hi], w.vel[lo:hi], w.next_waypoint[lo:hi], w.radius[lo:hi],
                              w.target_speed[lo:hi], w.mode[lo:hi])
        self.lo, self.hi = lo, hi
        if len(w.borders):
            self.ctx.set_borders(w.borders, w.section_center, w.section_length)
        if len(w.static_obstacles):
            self.ctx.set_obstacles(native.STATIC_OBSTACLE, [c for c, _ in w.static_obstacles],
                                   [r for _, r in w.static_obstacles])
        self.device_vehicles = bool(device_vehicles) and w.veh_center is not None and len(w.veh_center) > 0
        self._ticks = 0
        if self.device_vehicles:
            # the vehicle set lives on the device (replicated on every rank): centres advance ballistically and the
            # ellipse rings are regenerated there every tick (obstacles.py:269-281,297-329) -- no per-tick upload
            self.ctx.set_vehicles(w.veh_center, w.veh_yaw, w.veh_vel, w.veh_extent, w.veh_resolution)
        else:
            self.set_vehicles(w.vehicles_at(0))
        if self.world > 1 and self.exchange_mode == 'peer':
            self._bind_peers()
            self.ctx.stage()                               # collective: rows pushed to every rank + barrier
        else:
            self.ctx.stage()
            self._bind_gather()
            self.exchange()

    def load_lifecycle(self, w, life):
        """Mode machines and routes of this rank's rows on the device; ``tick`` then runs the whole
        SimulationRunner.tick sequence per call.  ``life`` is a ``synth.Lifecycle`` (or anything with its attributes)."""
        lo, hi = self.lo, self.hi
        idle = np.asarray(life.idle[lo:hi], dtype=bool)
        speed = w.target_speed[lo:hi]
        self.ctx.update_targets(mode=np.where(idle, 0, w.mode[lo:hi]).astype(np.uint8))
        self.ctx.set_mode_machines(speed, np.asarray(life.crossing_speed_factor[lo:hi]) * speed,
                                   life.crossing_safety_margin[lo:hi], np.where(idle, 0.0, speed),
                                   np.where(idle, 5.0, -1.0), 5.0)
        self.ctx.set_routes(life.routes[lo:hi], life.waypoint_threshold, fused=True)

    def tick(self, vehicles=None):
        """One headless SimulationRunner.tick: vehicles (the host 6-tuple, replicated on every rank, or the device-resident
        set advancing itself) -> mode machines + gap acceptance -> forces, velocities, hand-overs, positions.  Gap
        acceptance and the dynamic-obstacle force see the SAME vehicle state, as in the reference (run_simulation.py:92-102)."""
        sim_time = self._ticks * self.step_length
        if vehicles is not None:
            self.ctx.set_obstacles(native.DYNAMIC_OBSTACLE, vehicles[1], vehicles[5], vehicles[3])
            self.ctx.set_traffic(vehicles[1], vehicles[3], vehicles[4])
        self._advance_vehicles()
        self.ctx.tick_modes(sim_time)
        self._step_once(True)
        self._ticks += 1

    def _advance_vehicles(self):
        """Device-resident vehicles move at the top of a tick -- the simulator integrates the world before the SFM tick
        (run_simulation.py:77-95) -- and tick 0 sees the uploaded state, exactly like the host path ``vehicles_at(k)``."""
        if self.device_vehicles and self._ticks > 0:
            self.ctx.advance_vehicles(self.step_length)

    def set_vehicles(self, dyn_tuple):
        """The 6-tuple of pedestrian_simulation.py:108-115 (ids, centres, headings, velocities, extents, rings)."""
        if dyn_tuple is not None:
            self.ctx.set_obstacles(native.DYNAMIC_OBSTACLE, dyn_tuple[1], dyn_tuple[5], dyn_tuple[3])

    @property
    def peer(self):
        return self.world > 1 and self.exchange_mode == 'peer'

    def _bind_peers(self):
        """Exchange the CUDA IPC handles of (gather buffer, force accumulator, barrier flags) once."""
        mine = self.ctx.peer_export()
        table = [None] * self.world
        self.dist.all_gather_object(table, mine, group=self.group)
        self.ctx.peer_import(table)

    def _bind_gather(self):
        if self.world == 1:
            return
        ptr, per_rank = self.ctx.gather_buffer()
        view = _DeviceView(ptr, per_rank // 4 * self.world)
        self._gather = self.torch.as_tensor(view, device=f'cuda:{self.device}')
        self._per_rank = per_rank // 4
        ptr, per_rank = self.ctx.force_accumulator()
        view = _DeviceView(ptr, per_rank // 8 * self.world, '<i8')
        self._facc = self.torch.as_tensor(view, device=f'cuda:{self.device}')
        self._facc_per_rank = per_rank // 8

    def exchange(self):
        """All-gather every rank's staged block (in place: block r of the buffer is rank r's contribution)."""
        if self.world == 1:
            return
        mine = self._gather[self.rank * self._per_rank:(self.rank + 1) * self._per_rank]
        with self.torch.cuda.stream(self.stream):
            self.dist.all_gather_into_tensor(self._gather, mine, group=self.group)

    def reduce_forces(self):
        """Integer reduce-scatter of the fixed-point pair-force accumulators: every unordered pair was evaluated by
        exactly one rank, which added it to both pedestrians' rows; block r of the sum belongs to rank r (in place)."""
        mine = self._facc[self.rank * self._facc_per_rank:(self.rank + 1) * self._facc_per_rank]
        with self.torch.cuda.stream(self.stream):
            self.dist.reduce_scatter_tensor(mine, self._facc, group=self.group)

    # ---- stepping -------------------------------------------------------------------------------------------------
    def step(self, n_steps=1, integrate_positions=True):
        if self.device_vehicles:
            for _ in range(n_steps):
                self._advance_vehicles()
                self._step_once(integrate_positions)
                self._ticks += 1
            return
        self._ticks += n_steps
        if self.world == 1:
            self.ctx.step(n_steps, integrate_positions)
            return
        if self.peer:
            self.ctx.step_peer(n_steps, integrate_positions)
            return
        for _ in range(n_steps):
            self._step_once(integrate_positions)

    def _step_once(self, integrate_positions):
        if self.world == 1:
            self.ctx.step(1, integrate_positions)
        elif self.peer:
            self.ctx.step_peer(1, integrate_positions)
        else:
            self.ctx.step_begin()
            self.reduce_forces()
            self.ctx.step_end(integrate_positions)
            self.exchange()

    def tick_host(self, loc, vel, new_vel, new_loc=None):
        """One tick with host buffers for this rank's rows (H2D, kernels, D2H inside the call)."""
        self._advance_vehicles()
        self._ticks += 1
        if self.world == 1:
            self.ctx.tick_host(loc, vel, new_vel, new_loc)
            return
        self.ctx.update_kinematics(loc, vel)        # refresh -> restage -> exchange -> step, so every rank sees the
        self.ctx.stage()                            # other ranks' refreshed rows
        if self.peer:
            self.ctx.step_peer(1, new_loc is not None)
            self.ctx.download_state(new_loc if new_loc is not None else np.empty_like(new_vel), new_vel)
            return
        self.exchange()
        self.ctx.step_begin()
        self.reduce_forces()
        self.ctx.step_end(new_loc is not None)
        self.ctx.download_state(new_loc if new_loc is not None else np.empty_like(new_vel), new_vel)
        self.exchange()

    def local_state(self):
        return self.ctx.download_state()

    def synchronize(self):
        self.torch.cuda.synchronize(self.device)
        self.check_peers()

    def check_peers(self):
        """Raise if a flag barrier of the peer-memory exchange gave up waiting for another rank."""
        if not self.peer:
            return
        status = self.ctx.peer_status()
        if status['timed_out']:
            raise native.SfmError(f"peer-memory exchange: rank {self.rank} gave up waiting for rank(s) "
                                  f"{status['stalled_ranks']} at a flag barrier; results are invalid -- close every rank's "
                                  'engine and build new ones')

    def close(self):
        """Release the device context (and its mappings of the peers' buffers)."""
        self.ctx.close()
