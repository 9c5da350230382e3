#!/usr/bin/env python
"""Benchmark of the per-tick Social Force Model step (BASELINE.json metric) -- prints ONE JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3|cfg4|cfg5] [--peds N]

A step = one tick of the hot path over the whole synthetic crowd: all enabled force classes (all-pairs pedestrian force,
border / static-obstacle / dynamic-obstacle cell-list forces, acceleration force) + force sum + speed clamp + Euler
position/velocity update, followed on N > 1 ranks by the exchange of the staged rows.

  value        pair-interactions/s of the whole job = N (N - 1) K / t, t = sum of the K per-step device times (CUDA events
               on the launch stream, max over ranks), state resident in HBM.  agent-steps/s = value / (N - 1) beside it.
  e2e          the same metric through the host-buffer tick (sfm_tick_host): every step copies this rank's positions and
               velocities from pinned host memory to the device and the new positions / velocities back.
  e2e_dropin   (1 GPU) the reference's own plugin call, PedestrianSimulation.tick(sim_time) of the drop-in package, on the
               structured PedState table of the same crowd: table in, get_new_velocities() out.
  roofline     the all-pairs kernel (K1, timed alone at every world size) against the FP32 issue peak; roofline_k2 the
               cell-list kernels (distance evaluations/s + HBM bytes); roofline_hbm the integrate kernel (K3).
  parity       after the timed ticks: the total force of one more tick on >= 32 rows of the EVOLVED state (incl. the rows
               farthest from the origin) against the float64 oracle, and -- on N > 1 ranks -- the whole tick repeated on one
               GPU and compared bit for bit (forces, positions, velocities of every pedestrian).
  extra        the north-star configurations as short runs: cfg5 (N = 1,048,576; on 1 GPU also N = 370,688, the
               constant-pair-work point of the 1M weak ladder) and cfg4 (N = 262,144 x 2,048 device-resident vehicles).
  workload     1 GPU: BASELINE.json configs[2] (N = 65,536 + 1M border points + 50k obstacle points).  N GPUs: the
               constant-pair-work weak ladder anchored there, N_G = 65,536 sqrt(G) rounded to 256 G rows.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import time
import tomllib

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'carla-social-force-model_b200')
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np                     # noqa: E402

K1_INSTR_PER_PAIR = 58                 # FP32-pipe instructions per ordered pair (81 FLOP, 5 MUFU)
K1_INSTR_PER_PAIR_2D = 48              # same table: the 2-D / radius-off specialisation (what a flat crowd needs)
K1_INSTR_DOUBLE_SINGLE = 4             # + (hi_j - hi_i) + (lo_j - lo_i) instead of one subtraction, x and y
PLANE_BYTES = 60                       # staged float32 planes per pedestrian (csrc/sfm_common.cuh: 15 planes)
# K3 algorithmic bytes per agent-step: read loc+r 32, vel+speed 32, waypoint 16, pair force 24, three
# cell-list forces 3 x 16; write loc 32, vel 32, float32 staging planes 60, total force 24
K3_BYTES_PER_AGENT = 32 + 32 + 16 + 24 + 48 + 32 + 32 + PLANE_BYTES + 24
SMS, LANES = 148, 128
CFG5_N1 = 370688                       # N_8 / sqrt(8) rounded to 256: the 1-GPU point of the 1M constant-pair-work ladder


def load_config():
    with open(os.path.join(PKG, 'config', 'sfm_config.toml'), 'rb') as f:
        return tomllib.load(f)


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), 'measured'
    return {'hbm_gbs': 6650.0, 'sm_max_mhz': 1965.0}, 'fallback'


def ladder_n(world):
    """Constant pair work per GPU: N_G = 65,536 sqrt(G), rounded so that every rank owns whole 256-row tiles (then the
    multi-rank tick is bit-identical to the single-GPU tick of the same crowd, which `parity` checks)."""
    unit = 256 * world
    return max(unit, int(round(65536 * math.sqrt(world) / unit)) * unit)


def build_workload(workload, n, world):
    from sfm_b200 import synth
    if workload == 'cfg3':
        n = n or ladder_n(world)
        w = synth.make_config(3, n=n)
        name = 'cfg3: N=65,536 + 1,050,000 border points (5,000 sections) + 50,000 static-obstacle points' if n == 65536 \
            else f'cfg3 weak ladder: N={n} (65,536*sqrt({world})), border/obstacle sets scaled by area'
    elif workload == 'cfg4':
        w = synth.make_config(4, n=n)
        name = f'cfg4: N={w.n} x {len(w.veh_center)} vehicles (68-point rings, cutoff 50 m)'
    elif workload == 'cfg5':
        w = synth.make_config(5, n=n)
        name = f'cfg5: N={w.n} synthetic city-scale crowd'
    else:
        raise SystemExit(f'unknown workload {workload}')
    return w, name


def bench_config(name, w, world, workload):
    """The `config` object -- identical in both arms (the driver compares them)."""
    return {'workload': name, 'n_pedestrians': w.n, 'rows_per_gpu': int(math.ceil(w.n / world)), 'step_length': w.step_length,
            'forces': 'all five on',
            'l2': 'GPU arm: L2 flushed between timed steps (256 MiB fill, outside the timed spans); CPU arm: n/a',
            'partition': f'row blocks over {world} rank(s)',
            'staged_order': 'GPU arm: each rank stages its rows along a Hilbert curve, rebuilt on the device every 32 ticks '
                            '(speed only; row order and results per row unchanged); CPU arm: n/a',
            'vehicles': ('device-resident: centres advanced and ellipse rings regenerated on the device every tick'
                         if workload == 'cfg4' else 'none' if w.veh_center is None else 'host rings')}


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML from a background thread, 10 ms period)."""
    REASONS = {0x8: 'hw_slowdown', 0x40: 'hw_thermal_slowdown', 0x20: 'sw_thermal_slowdown', 0x4: 'sw_power_cap'}

    def __init__(self, device):
        import threading
        self.samples, self.reasons, self.power, self.max_mhz = [], set(), [], None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # torch's device index follows CUDA_VISIBLE_DEVICES; NVML's does not
            visible = os.environ.get('CUDA_VISIBLE_DEVICES')
            index = int(visible.split(',')[device]) if visible and visible.split(',')[device].isdigit() else device
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self._nvml = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        except Exception:
            self._thread = None

    def _run(self):
        nv = self._nvml
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self._h) / 1000.0)
            except Exception:
                pass
            self._stop.wait(0.01)

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': [], 'samples': 0}
        if self._thread is None:
            return out
        self._stop.set()
        self._thread.join(timeout=2)
        if self.samples:
            out.update(sm_mhz=float(np.median(self.samples)), reasons=sorted(self.reasons), samples=len(self.samples),
                       power_w_max=max(self.power) if self.power else None)
        return out


def run_reference(args, world, rank):
    """The reference arm: the path's CPU implementation on this box's host cores (oracle port, all cores)."""
    if rank != 0:
        return
    from oracle import cpu_baseline
    cfg = load_config()
    w, name = build_workload(args.workload, args.n, world)
    cores = os.cpu_count() or 1
    # K + W samples of ~2-3 s each, shrunk for long runs so that the whole arm stays within a few minutes
    rows_per_core = max(1, int(args.cpu_rows_per_core // 4 * min(1.0, 20.0 / max(args.steps, 1))))
    for _ in range(args.warmup):
        cpu_baseline.time_sample(w, cfg, rows_per_core=max(1, rows_per_core // 8), cores=cores)
    times, rows = [], 0
    for k in range(args.steps):
        r = cpu_baseline.time_sample(w, cfg, rows_per_core=rows_per_core, cores=cores, seed=k)
        times.append(r['seconds'])
        rows, cores = r['rows'], r['cores']           # the workers actually used (memory-capped on many-core boxes)
    sec = float(np.mean(times))
    value = rows * (w.n - 1) / sec
    sample = f'{rows} of {w.n} rows per step (full tick for those rows against all {w.n} pedestrians), {cores} forked workers'
    line = {
        'impl': 'reference', 'metric': 'pair_interactions_per_s', 'value': value, 'unit': 'pair-interactions/s',
        'agent_steps_per_s': value / (w.n - 1), 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': sec * 1e3 * w.n / rows, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': bench_config(name, w, world, args.workload),
        'note': 'ms_per_step extrapolated from the row sample (the float64 numpy port cannot finish a whole tick in minutes)',
        'cpu_baseline': {'value': value, 'unit': 'pair-interactions/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'pair-interactions/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


# ---- helpers of the GPU arm --------------------------------------------------------------------------------------------
class Job:
    """torch / torch.distributed plumbing shared by the measurements of one bench run."""

    def __init__(self, world, rank, local_rank):
        import torch
        self.torch, self.world, self.rank, self.local_rank = torch, world, rank, local_rank
        self.dist = torch.distributed
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')       # > 126 MB L2

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device='cuda')
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def gather(self, obj):
        """Python object from every rank, in rank order, on every rank."""
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out


def timed_ticks(job, e, steps, flush=True):
    """K ticks, each bracketed by CUDA events on the launch stream (L2 flushed before each, outside the span); returns
    (sum of per-tick ms as max over ranks, per-tick list of this rank, wall seconds)."""
    torch = job.torch
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    job.barrier()
    t_wall = time.perf_counter()
    for k in range(steps):
        if flush:
            with torch.cuda.stream(e.stream):
                job.flush.fill_(k & 0xff)
        starts[k].record(e.stream)
        e.step(1, True)
        stops[k].record(e.stream)
    job.barrier()
    wall = time.perf_counter() - t_wall
    per_step = [a.elapsed_time(b) for a, b in zip(starts, stops)]
    return job.max_over_ranks([sum(per_step)])[0], per_step, wall


def sample_rows(loc, n_uniform, n_far):
    centre = np.round((loc[:, :2].min(axis=0) + loc[:, :2].max(axis=0)) * 0.5)
    far = np.argsort(-np.abs(loc[:, :2] - centre).max(axis=1))[:n_far]
    return np.unique(np.concatenate([np.linspace(0, len(loc) - 1, n_uniform).astype(np.int64), far]))


def parity_check(job, e, w, cfg, n_uniform=24, n_far=8, chunk=8):
    """One more tick on the evolved state, checked two ways (module docstring).  Collective: every rank takes part; rank 0
    returns the report."""
    from sfm_b200 import native
    from sfm_b200.engine import crowd_origin
    loc_l, vel_l = e.local_state()
    order_l = e.ctx.slot_order()                   # the staged order this tick's pair kernel reads its rows in
    e.step(1, True)
    f_l = e.ctx.download_force()
    loc2_l, vel2_l = e.local_state()
    e.synchronize()
    parts = job.gather((loc_l, vel_l, f_l, loc2_l, vel2_l, order_l + int(e.lo)))
    if job.rank != 0:
        job.barrier()                              # wait on the host while rank 0 checks (no device-side spinning)
        return None
    loc, vel, force, loc2, vel2, order = (np.concatenate([p[k] for p in parts]) for k in range(6))
    report = {'tick': 'the tick after the timed and host-buffer ticks (evolved, non-float32-exact state)'}
    # (1) float64 oracle on a row sample: forces.py on the float64 state (1e-4 rel + 1e-5 abs per component, plus the
    #     magnitude carried by pairs within 1e-5 rad of the model's sign / wrap discontinuities)
    from oracle import sfm_oracle as O
    rows = sample_rows(loc, n_uniform, n_far)
    scene = O.Scene(cfg, w.step_length, w.borders, w.section_center, w.section_length, w.static_obstacles)
    t0 = time.perf_counter()
    if w.veh_center is None:                       # (with device-resident vehicles only the bitwise check below applies)
        per_class = O.forces_by_class(scene, loc, vel, w.next_waypoint, w.radius, w.target_speed, w.mode, rows=rows, chunk=chunk)
        want = O.total_force(per_class, len(rows))
        got = force[rows]
        what = 'total force (all classes)'
        _, risk = O.pedestrian_force(loc, vel, w.radius, scene.ped, scene.use_ped_radius, rows=rows, chunk=chunk,
                                     return_risk=True)
        err = np.abs(got - want)
        tol = 1e-5 + 1e-4 * np.abs(want) + risk[:, None]
        report['oracle'] = {'rows': int(len(rows)), 'rows_farthest_from_origin': int(n_far), 'quantity': what,
                            'tolerance': '1e-4 rel + 1e-5 abs per component (+ discontinuity risk)',
                            'worst_err_over_tol': float((err / tol).max()), 'max_abs_err': float(err.max()),
                            'ok': bool((err <= tol).all()), 'seconds': time.perf_counter() - t0}
    # (2) the same tick on ONE GPU, bit for bit (integer accumulation makes the pair force independent of the partition)
    if job.world > 1:
        ctx = native.Context(job.local_rank)
        ctx.set_params(native.params_from_config(cfg, w.step_length))
        ctx.set_origin(*crowd_origin(w.loc))
        ctx.upload_state(loc, vel, w.next_waypoint, w.radius, w.target_speed, w.mode)
        # same staged order as the ranks (each rank's own order, shifted to its row block): the float32 tile partials of
        # the pair force follow the tile composition, everything above them is integer or float64
        ctx.set_slot_order(order)
        if len(w.borders):
            ctx.set_borders(w.borders, w.section_center, w.section_length)
        if len(w.static_obstacles):
            ctx.set_obstacles(native.STATIC_OBSTACLE, [c for c, _ in w.static_obstacles], [r for _, r in w.static_obstacles])
        ctx.step(1, True)
        f1 = ctx.download_force()
        l1, v1 = ctx.download_state()
        ctx.close()
        same = bool(np.array_equal(f1, force) and np.array_equal(l1, loc2) and np.array_equal(v1, vel2))
        checksum = lambda a: hex(int(np.ascontiguousarray(a).view(np.uint64).sum(dtype=np.uint64)))     # noqa: E731
        report['single_gpu_bitwise'] = {
            'identical': same, 'pedestrians': int(len(loc)), 'arrays': 'total force, new positions, new velocities',
            'max_abs_force_diff': float(np.abs(f1 - force).max()),
            'force_checksum': {'ranks': checksum(force), 'single_gpu': checksum(f1)}}
    report['ok'] = bool(report.get('oracle', {}).get('ok', True) and report.get('single_gpu_bitwise', {}).get('identical', True))
    job.barrier()
    return report


def short_run(job, cfg, workload, n, ticks=3, with_parity=False):
    """One north-star configuration as a short run: 1 untimed + `ticks` timed ticks (+ parity)."""
    from sfm_b200 import engine as eng
    w, name = build_workload(workload, n, job.world)
    e = eng.Engine(cfg, w.step_length, device=job.local_rank)
    e.load(w, device_vehicles=(workload == 'cfg4'))
    e.step(1 + (2 if job.world > 1 else 0), True)
    ms_total, _, _ = timed_ticks(job, e, ticks)
    e.check_peers()
    ms = ms_total / ticks
    out = {'workload': name, 'n_pedestrians': w.n, 'n_gpus': job.world, 'ticks': ticks, 'ms_per_step': ms,
           'pair_interactions_per_s': w.n * (w.n - 1) / (ms * 1e-3), 'agent_steps_per_s': w.n / (ms * 1e-3)}
    if with_parity:
        rep = parity_check(job, e, w, cfg, n_uniform=8, n_far=8, chunk=2)
        if rep is not None:
            out['parity'] = rep
    e.ctx.close()
    return out


def dropin_ticks(cfg, w, ticks, warmup):
    """PedestrianSimulation.tick(sim_time) on the structured PedState table (the reference's own plugin API), CARLA
    stubbed: table in, get_new_velocities() out, positions advanced by the caller like the simulator would."""
    import pedestrian_simulation
    from ped_mode_manager import PedMode, PedModeManager
    from sfm_b200.session import reset_session
    reset_session()
    sim = pedestrian_simulation.PedestrianSimulation(list(w.borders), w.section_info(), list(w.static_obstacles), cfg,
                                                     w.step_length, record_states=False)
    names = [f'p_{i}' for i in range(w.n)]
    modes = [PedModeManager(names[i], float(w.target_speed[i]), PedMode(int(w.mode[i])), 1.5, 1.0) for i in range(w.n)]
    sim.peds.add_pedestrians(names, np.arange(w.n), w.loc, w.vel, w.next_waypoint, modes, w.radius, w.target_speed)
    times = []
    for k in range(warmup + ticks):
        t0 = time.perf_counter()
        sim.tick(k * w.step_length)
        nv = sim.get_new_velocities()
        dt = time.perf_counter() - t0
        if k >= warmup:
            times.append(dt * 1e3)
        sim.peds.state['loc'] += nv['vel'] * w.step_length
    reset_session()
    return {'ms_per_step': float(np.mean(times)), 'ms_min': float(min(times)), 'ticks': ticks,
            'value': w.n * (w.n - 1) / (float(np.mean(times)) * 1e-3), 'unit': 'pair-interactions/s',
            'api': 'PedestrianSimulation.tick(sim_time) + get_new_velocities() on the structured PedState table '
                   '(pedestrian_simulation.py:57-83), record_states=False, stock PedModeManager objects',
            'h2d_bytes_per_step': int(w.n * sim.peds.state.dtype.itemsize), 'd2h_bytes_per_step': int(w.n * 32)}


def run_ours(args, world, rank, local_rank):
    import torch
    from sfm_b200 import engine as eng, native
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm')
    torch.cuda.set_device(local_rank)
    dist = torch.distributed
    if world > 1:
        # NCCL only carries the set-up traffic (handle table, barriers, the gathers of `parity`) unless SFM_EXCHANGE=nccl;
        # its INIT lines (rank / nranks / transport) go to stderr so the communicator is observable, stdout stays one line
        os.environ['NCCL_DEBUG'] = os.environ.get('SFM_NCCL_DEBUG', 'INFO')     # (an inherited VERSION / WARN level would
        os.environ['NCCL_DEBUG_SUBSYS'] = os.environ.get('SFM_NCCL_DEBUG_SUBSYS', 'INIT')     #  print its banner on stdout)
        os.environ['NCCL_DEBUG_FILE'] = '/dev/stderr'
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    job = Job(world, rank, local_rank)
    cfg = load_config()
    w, name = build_workload(args.workload, args.n, world)
    n = w.n
    e = eng.Engine(cfg, w.step_length, device=local_rank)
    e.load(w, device_vehicles=(args.workload == 'cfg4'))     # cfg4: vehicle rings regenerated on the device every tick
    ctx = e.ctx
    stream = e.stream

    # ---- device-resident timing ---------------------------------------------------------------------------------
    for _ in range(args.warmup + (3 if world > 1 else 0)):     # NCCL builds its channels lazily: a few extra untimed steps
        e.step(1, True)
    sampler = ClockSampler(local_rank) if rank == 0 else None            # NVML start-up costs ~0.1 s on rank 0 only:
    ctx.reset_stats()                                                     # keep it in front of the barrier, or the other
    ctx.set_profiling(True)                                               # ranks' first timed tick waits for rank 0
    ms_total, per_step, wall = timed_ticks(job, e, args.steps)
    clocks = sampler.stop() if sampler else None
    stats = ctx.stats()
    ctx.set_profiling(False)
    ms_pairs, ms_integrate, ms_sets = job.max_over_ranks([stats['ms_pairs'], stats['ms_integrate'],
                                                          stats['ms_segments'] + stats['ms_cells']])
    ms_per_step = ms_total / args.steps
    value = n * (n - 1) / (ms_per_step * 1e-3)

    # ---- end-to-end through host buffers --------------------------------------------------------------------------
    rows = e.hi - e.lo
    pin = lambda: torch.empty((max(rows, 1), 3), dtype=torch.float64).pin_memory()      # noqa: E731
    h_loc, h_vel, h_nloc, h_nvel = pin(), pin(), pin(), pin()
    loc0, vel0 = e.local_state()
    h_loc.numpy()[:rows], h_vel.numpy()[:rows] = loc0, vel0
    a_loc, a_vel, a_nloc, a_nvel = (x.numpy()[:rows] for x in (h_loc, h_vel, h_nloc, h_nvel))
    e2e_steps = max(3, min(args.steps, 20))
    e2e_ms = 0.0
    for k in range(args.warmup + e2e_steps):
        with torch.cuda.stream(stream):
            job.flush.fill_(k & 0xff)
        job.barrier()
        t0 = time.perf_counter()
        e.tick_host(a_loc, a_vel, a_nvel, a_nloc)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if k >= args.warmup:
            e2e_ms += dt * 1e3
        a_loc[...] = a_nloc                       # next tick's input = this tick's output (the co-simulation loop)
        a_vel[...] = a_nvel
    e2e_ms_per_step = job.max_over_ranks([e2e_ms])[0] / e2e_steps
    e2e_value = n * (n - 1) / (e2e_ms_per_step * 1e-3)
    e.check_peers()                               # a timed-out flag barrier would have produced garbage: fail loudly

    # ---- parity on the evolved state (collective) --------------------------------------------------------------------
    parity = parity_check(job, e, w, cfg) if not args.no_parity else None

    # ---- kernels alone: the cell-list kernels, then the pair kernel (no cell-list kernel beside it), every world size ----
    k2 = {'ms': 0.0, 'pairs': 0, 'evals': 0, 'classes': []}
    scratch = np.empty((rows, 3))
    for cls, label in ((native.BORDER, 'border'), (native.STATIC_OBSTACLE, 'static'), (native.DYNAMIC_OBSTACLE, 'dynamic')):
        present = (len(w.borders) > 0) if cls == native.BORDER else (len(w.static_obstacles) > 0) \
            if cls == native.STATIC_OBSTACLE else e.device_vehicles or w.veh_center is not None
        if not present:
            continue
        ctx.force(cls, scratch)
        ctx.reset_stats()
        ctx.set_profiling(True)
        for _ in range(5):
            with torch.cuda.stream(stream):
                job.flush.fill_(2)
            ctx.force(cls, scratch)
        s2 = ctx.stats()
        ctx.set_profiling(False)
        pairs, evals = ctx.count_point_evaluations(cls)
        k2['ms'] += (s2['ms_segments'] + s2['ms_cells']) / 5
        k2['pairs'] += pairs
        k2['evals'] += evals
        k2['classes'].append(label)
    only_pairs = native.params_from_config(cfg, w.step_length, enable={'pedestrian_force': True, 'acceleration_force': True})
    ctx.set_params(only_pairs)
    e.step(2, True)
    ctx.reset_stats()
    ctx.set_profiling(True)
    for _ in range(5):
        with torch.cuda.stream(stream):
            job.flush.fill_(1)
        e.step(1, True)
    iso = ctx.stats()
    ctx.set_profiling(False)
    ctx.set_params(native.params_from_config(cfg, w.step_length))
    k1_ms, k2_ms = job.max_over_ranks([iso['ms_pairs'] / max(iso['pair_launches'], 1), k2['ms']])
    k1_evals = iso['pair_evaluations'] / max(iso['pair_launches'], 1)
    tile_pairs = lambda st: st['pair_evaluations'] / 65536.0                                           # noqa: E731
    local_frac_alone = iso['local_tile_pairs'] / max(tile_pairs(iso), 1.0)
    local_frac_timed = stats['local_tile_pairs'] / max(tile_pairs(stats), 1.0)
    k2_pairs, k2_evals = (int(sum(v)) for v in zip(*job.gather((k2['pairs'], k2['evals']))))
    e.check_peers()
    e.ctx.close()

    # ---- the north-star configurations as short runs (collective) ---------------------------------------------------
    extra = {}
    if args.workload == 'cfg3' and not args.no_extra:
        extra['cfg5'] = short_run(job, cfg, 'cfg5', None, ticks=3, with_parity=world > 1)
        if world == 1:
            extra['cfg5_n1'] = short_run(job, cfg, 'cfg5', CFG5_N1, ticks=3)
            extra['cfg5_n1']['note'] = ('1-GPU point of the 1M constant-pair-work weak ladder: efficiency(8) = rate(8 GPUs, '
                                        'N=1,048,576) / (8 x this rate)')
        extra['cfg4'] = short_run(job, cfg, 'cfg4', None, ticks=3)

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        sm_max_hz = float(peaks.get('sm_max_mhz', 1965.0)) * 1e6
        fp32_peak = SMS * LANES * sm_max_hz / 1e12                      # T FP32-pipe lane-instructions / s
        k1_ms_in_step = ms_pairs / max(stats['pair_launches'], 1)       # overlapped with the cell-list kernels
        # pair terms the launch actually evaluates (the symmetric kernel visits every unordered pair once and applies it
        # to both rows; padded slots included) -- this is what occupies the pipes -- and the ordered pairs it covers
        flat = bool(np.all(w.loc[:, 2] == w.loc[0, 2]) and not np.any(w.vel[:, 2])) and not cfg.get('use_ped_radius', False)
        instr = K1_INSTR_PER_PAIR_2D if flat else K1_INSTR_PER_PAIR
        k1_rate = k1_evals / (k1_ms * 1e-3)
        k1_ordered_rate = (e.hi - e.lo) * (n - 1) / (k1_ms * 1e-3)
        achieved = k1_rate * instr / 1e12
        traffic, traffic_source = None, 'no ncu capture for this (N, world): null'
        prof = os.path.join(ROOT, 'profiles', 'k1_ncu_summary.json')          # ncu --set full capture of k1_sym_pairs
        if os.path.exists(prof):
            with open(prof) as f:
                cap = json.load(f)
            if cap.get('n_pedestrians') == n and cap.get('world', 1) == world:
                traffic = cap.get('dram_bytes_per_launch')
                traffic_source = f"static: ncu --set full capture at N={n}, 1 rank ({cap.get('source', 'profiles/')}); not re-measured per run"
        k3_ms = ms_integrate / max(stats['steps'], 1)
        k3_gbs = rows * K3_BYTES_PER_AGENT / (k3_ms * 1e-3) / 1e9 if k3_ms > 0 else None
        hbm = float(peaks.get('hbm_gbs', 6650.0))
        n_points = sum(len(b) for b in w.borders) + sum(len(r) for _, r in w.static_obstacles)
        n_items = len(w.borders) + len(w.static_obstacles)
        k2_bytes = 40 * rows + 8 * n_points + 20 * n_items             # per step, this rank
        transport = 'none (1 rank)' if world == 1 else \
            ('peer memory over NVLink (CUDA IPC): reduce-scatter fused into k1_sym_finish, all-gather into k3_integrate, '
             'flag barriers (K7); NCCL carries set-up and verification traffic only' if e.peer
             else 'NCCL reduce_scatter_tensor + all_gather_into_tensor')
        line = {
            'metric': 'pair_interactions_per_s', 'value': value, 'unit': 'pair-interactions/s',
            'agent_steps_per_s': value / (n - 1), 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32 pair forces (double-single positions on neighbouring tiles, run-local float32 positions in the far field) / '
                     'f64 state, cell-list forces and integration',
            'data': 'synthetic', 'config': bench_config(name, w, world, args.workload), 'transport': transport,
            'exchange_bytes_per_pedestrian': {'all_gather': PLANE_BYTES, 'reduce_scatter': 32},
            'e2e': {'value': e2e_value, 'unit': 'pair-interactions/s', 'ms_per_step': e2e_ms_per_step,
                    'h2d_bytes_per_step': int(rows * 48), 'd2h_bytes_per_step': int(rows * 48), 'steps': e2e_steps,
                    'api': 'sfm_tick_host (pinned host loc/vel in, new loc/vel out)'},
            'gpu_launches': int(stats['launches']),
            'kernel_ms_per_step': {'pairs_k1': ms_pairs / args.steps, 'segments_cells_k2': ms_sets / args.steps,
                                   'integrate_k3': ms_integrate / args.steps},
            'roofline': {'bound': 'fp32_issue', 'achieved': achieved, 'peak': fp32_peak,
                         'unit': 'T FP32-pipe instr/s', 'frac': achieved / fp32_peak, 'traffic': traffic,
                         'traffic_source': traffic_source,
                         'kernel': 'k1_sym_pairs (+ accumulator zero/finish), timed alone on this world size',
                         'ms_per_launch': k1_ms,
                         'pair_terms_evaluated_per_s': k1_rate, 'ordered_pairs_covered_per_s': k1_ordered_rate,
                         'frac_ordered_equivalent': k1_ordered_rate * instr / 1e12 / fp32_peak,
                         'frac_with_double_single_differences':
                             k1_rate * (instr + K1_INSTR_DOUBLE_SINGLE * (1.0 - local_frac_alone)) / 1e12 / fp32_peak,
                         'ms_per_launch_inside_step': k1_ms_in_step, 'instr_per_pair': instr,
                         'local_tile_pair_fraction': {'timed_ticks': local_frac_timed, 'alone': local_frac_alone,
                                                      'meaning': 'share of the 256 x 256 tile pairs (rank 0) read through '
                                                                 'run-local origins: one subtraction per coordinate; the '
                                                                 'rest -- neighbouring tiles -- take the double-single form'},
                         'note': 'f_ji = -f_ij exactly, so each unordered pair is evaluated once: achieved/frac count '
                                 'the evaluated pair terms (hardware utilisation); the ordered-pair figures are the '
                                 'useful work the metric counts.  frac uses SURVEY 8d\'s per-pair figure; on neighbouring '
                                 'tiles the kernel additionally forms d = (hi_j - hi_i) + (lo_j - lo_i) (+4 instr per pair), '
                                 'which the 1e-4 / 1e-5 parity bar needs on evolved states -- frac_with_double_single_'
                                 'differences weighs that by the share of tile pairs that took it',
                         'algorithmic': (f'{instr} FP32-pipe instr per pair term (SURVEY 8d: 58 = 3-D radius-on, 81 FLOP, '
                                         '5 MUFU; 48 = the 2-D radius-off specialisation a flat crowd takes)'),
                         'peak_source': f'148 SM x 128 lanes x sm_max_mhz ({peak_kind} MEASURED_PEAKS.json clock); '
                                        'FFMA microbenchmark on this pool: 33.2 T/s (profiles/microbench)'},
            'roofline_k2': {'bound': 'hbm', 'kernel': 'k2_segments<border|obstacle> + pedestrian binning, timed alone: ' + '+'.join(k2['classes']),
                            'ms_alone': k2_ms, 'algorithmic_bytes': int(k2_bytes),
                            'achieved': (k2_bytes / (k2_ms * 1e-3) / 1e9) if k2_ms > 0 else None, 'peak': hbm, 'unit': 'GB/s',
                            'frac': (k2_bytes / (k2_ms * 1e-3) / 1e9 / hbm) if k2_ms > 0 else None,
                            'pedestrian_item_pairs_inside_cutoff': k2_pairs,
                            'reference_distance_evaluations': k2_evals,
                            'reference_distance_evaluations_per_s': (k2_evals / (k2_ms * 1e-3)) if k2_ms > 0 else None,
                            'note': 'SURVEY 8d names HBM as the bound; the kernels are latency / issue bound far below it '
                                    '(DESIGN.md K2), so the distance-evaluation rate is the figure to follow: the distances '
                                    'the reference\'s argmin ranges over (forces.py:154,228) per second of kernel time; the '
                                    'kernels evaluate ~1-2 of them per pair in float64 (chord-projection window)',
                            'ncu_static': 'profiles/r2_ncu_k2_*.csv (issue %, dram__bytes)'},
            'roofline_hbm': {'bound': 'hbm', 'kernel': 'k3_integrate', 'achieved': k3_gbs, 'peak': hbm, 'unit': 'GB/s',
                             'frac': (k3_gbs / hbm) if k3_gbs else None, 'ms_per_launch': k3_ms, 'peak_kind': peak_kind},
            'parity': parity, 'extra': extra,
            'clocks': clocks, 'wall_s_timed_loop': wall,
            'ms_per_step_rank0': {'median': float(np.median(per_step)), 'min': float(min(per_step)), 'max': float(max(per_step))},
        }
        if clocks and clocks.get('sm_mhz'):
            line['roofline']['frac_at_sampled_clock'] = achieved / (SMS * LANES * clocks['sm_mhz'] * 1e6 / 1e12)
        # the two rank-0-only legs below must never cost the main line: a failure is recorded, not raised
        if world == 1 and args.workload == 'cfg3' and not args.no_dropin:
            try:
                line['e2e_dropin'] = dropin_ticks(cfg, w, ticks=max(3, min(args.steps, 20)), warmup=3)
                line['e2e_dropin']['ratio_to_sfm_tick_host'] = line['e2e_dropin']['ms_per_step'] / e2e_ms_per_step
            except Exception as err:                                    # noqa: BLE001
                line['e2e_dropin'] = {'error': f'{type(err).__name__}: {err}'}
        if world == 1 and not args.no_cpu_baseline:
            try:
                from oracle import cpu_baseline
                cores = os.cpu_count() or 1
                r = cpu_baseline.time_sample(w, cfg, rows_per_core=args.cpu_rows_per_core, cores=cores)
                line['cpu_baseline'] = {
                    'value': r['pairs_per_s'], 'unit': 'pair-interactions/s', 'cores': r['cores'], 'kind': 'port',
                    'agent_steps_per_s': r['agent_steps_per_s'], 'seconds': r['seconds'],
                    'sample': f"one full tick for {r['rows']} of {n} rows (each against all {n} pedestrians + the border/"
                              f"obstacle sets), float64 numpy oracle, {r['cores']} forked workers"}
            except Exception as err:                                    # noqa: BLE001
                line['cpu_baseline'] = {'error': f'{type(err).__name__}: {err}'}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cfg3', choices=['cfg3', 'cfg4', 'cfg5'])
    ap.add_argument('--peds', dest='n', type=int, default=None, help='override the pedestrian count')
    ap.add_argument('--cpu-rows-per-core', type=int, default=384,
                    help='rows of the CPU sample per host core (cpu_baseline leg: ~10 s; the reference arm uses a quarter per step)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-parity', action='store_true')
    ap.add_argument('--no-extra', action='store_true', help='skip the cfg5 / cfg4 short runs')
    ap.add_argument('--no-dropin', action='store_true')
    args = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus and world > 1:
        raise SystemExit(f'--gpus {args.gpus} but WORLD_SIZE={world}')
    if args.gpus > 1 and 'WORLD_SIZE' not in os.environ and args.impl == 'ours':
        import socket                     # launched without torchrun: start one rank per GPU ourselves
        with socket.socket() as sk:
            sk.bind(('127.0.0.1', 0))
            port = sk.getsockname()[1]
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={args.gpus}',
               '--master-addr', '127.0.0.1', '--master-port', str(port), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == 'reference':
        run_reference(args, args.gpus, rank)
    else:
        args.warmup = max(args.warmup, 3)
        run_ours(args, world, rank, local_rank)


if __name__ == '__main__':
    main()
