#!/usr/bin/env python
"""Benchmark of the per-tick Social Force Model step (BASELINE.json metric) -- prints ONE JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3|cfg4|cfg5] [--peds N]

A step = one tick of the hot path over the whole synthetic crowd: all enabled force classes (all-pairs pedestrian force,
border / static-obstacle / dynamic-obstacle cell-list forces, acceleration force) + force sum + speed clamp + Euler
position/velocity update, followed on N > 1 ranks by the all-gather of the staged rows.

  value     pair-interactions/s of the whole job = N (N - 1) K / t, t = sum of the K per-step device times (CUDA events
            on the launch stream, max over ranks), state resident in HBM.  agent-steps/s = value / (N - 1) is reported
            beside it (`agent_steps_per_s`).
  e2e       the same metric through the host-buffer tick (sfm_tick_host): every step copies this rank's positions and
            velocities from pinned host memory to the device and the new positions / velocities back.
  roofline  the all-pairs kernel (K1) against the FP32 issue peak, plus the HBM-bound integrate kernel (K3) in `roofline_hbm`.
  workload  1 GPU: BASELINE.json configs[2] (N = 65,536 + 1M border points + 50k obstacle points).  N GPUs: the
            constant-pair-work weak ladder anchored there, N_G = 65,536 sqrt(G) (sets scaled by area).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
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
# K3 algorithmic bytes per agent-step: read loc+r 32, vel+speed 32, waypoint 16, pair force 12, three
# cell-list forces 3 x 16; write loc 32, vel 32, float32 staging planes 28, total force 24
K3_BYTES_PER_AGENT = 32 + 32 + 16 + 12 + 48 + 32 + 32 + 28 + 24
SMS, LANES = 148, 128


def load_config():
    with open(os.path.join(PKG, 'config', 'sfm_config.toml'), 'rb') as f:
        return tomllib.load(f)


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), 'measured'
    return {'hbm_gbs': 6650.0, 'sm_max_mhz': 1965.0}, 'fallback'


def build_workload(args, world):
    from sfm_b200 import synth
    if args.workload == 'cfg3':
        n = args.n or int(round(65536 * math.sqrt(world) / 256.0)) * 256
        w = synth.make_config(3, n=n)
        name = 'cfg3: N=65,536 + 1,050,000 border points (5,000 sections) + 50,000 static-obstacle points' if n == 65536 \
            else f'cfg3 weak ladder: N={n} (65,536*sqrt({world})), border/obstacle sets scaled by area'
    elif args.workload == 'cfg4':
        w = synth.make_config(4, n=args.n)
        name = f'cfg4: N={w.n} x {len(w.veh_center)} vehicles (68-point rings, cutoff 50 m)'
    elif args.workload == 'cfg5':
        w = synth.make_config(5, n=args.n)
        name = f'cfg5: N={w.n} synthetic city-scale crowd'
    else:
        raise SystemExit(f'unknown workload {args.workload}')
    return w, name


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
    w, name = build_workload(args, world)
    cores = os.cpu_count() or 1
    rows_per_core = max(1, args.cpu_rows_per_core // 4)         # K + W samples of ~2-3 s each
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
        'config': {'workload': name, 'n_pedestrians': w.n, 'note': 'ms_per_step extrapolated from the row sample'},
        'cpu_baseline': {'value': value, 'unit': 'pair-interactions/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'pair-interactions/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def run_ours(args, world, rank, local_rank):
    import torch
    from sfm_b200 import engine as eng
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm')
    torch.cuda.set_device(local_rank)
    dist = torch.distributed
    if world > 1:
        os.environ['NCCL_DEBUG'] = os.environ.get('SFM_NCCL_DEBUG', 'WARN')   # keep stdout to the one JSON line:
        os.environ['NCCL_DEBUG_FILE'] = '/dev/stderr'                           # NCCL's version banner goes to stderr
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    cfg = load_config()
    w, name = build_workload(args, world)
    n = w.n
    e = eng.Engine(cfg, w.step_length, device=local_rank)
    e.load(w, device_vehicles=(args.workload == 'cfg4'))     # cfg4: vehicle rings regenerated on the device every tick
    ctx = e.ctx
    stream = e.stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')       # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timing ---------------------------------------------------------------------------------
    for _ in range(args.warmup + (3 if world > 1 else 0)):     # NCCL builds its channels lazily: a few extra untimed steps
        e.step(1, True)
    sampler = ClockSampler(local_rank) if rank == 0 else None            # NVML start-up costs ~0.1 s on rank 0 only:
    ctx.reset_stats()                                                     # keep it in front of the barrier, or the other
    ctx.set_profiling(True)                                               # ranks' first timed tick waits for rank 0
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    t_wall = time.perf_counter()
    for k in range(args.steps):
        with torch.cuda.stream(stream):
            flush.fill_(k & 0xff)                                         # L2 flush, outside the timed span
        starts[k].record(stream)
        e.step(1, True)
        stops[k].record(stream)
    barrier()
    wall = time.perf_counter() - t_wall
    clocks = sampler.stop() if sampler else None
    per_step = [a.elapsed_time(b) for a, b in zip(starts, stops)]
    ms_total = sum(per_step)
    stats = ctx.stats()
    ctx.set_profiling(False)
    t = torch.tensor([ms_total, stats['ms_pairs'], stats['ms_integrate'], stats['ms_segments'] + stats['ms_cells']],
                     dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_pairs, ms_integrate, ms_sets = t.tolist()
    ms_per_step = ms_total / args.steps
    value = n * (n - 1) / (ms_per_step * 1e-3)

    # ---- end-to-end through host buffers --------------------------------------------------------------------------
    rows = e.hi - e.lo
    pin = lambda: torch.empty((max(rows, 1), 3), dtype=torch.float64).pin_memory()      # noqa: E731
    h_loc, h_vel, h_nloc, h_nvel = pin(), pin(), pin(), pin()
    loc0, vel0 = e.local_state()
    h_loc.numpy()[:rows], h_vel.numpy()[:rows] = loc0, vel0
    a_loc, a_vel, a_nloc, a_nvel = (x.numpy()[:rows] for x in (h_loc, h_vel, h_nloc, h_nvel))
    e2e_steps = max(3, min(args.steps, 10))
    e2e_ms = 0.0
    for k in range(args.warmup + e2e_steps):
        with torch.cuda.stream(stream):
            flush.fill_(k & 0xff)
        barrier()
        t0 = time.perf_counter()
        e.tick_host(a_loc, a_vel, a_nvel, a_nloc)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if k >= args.warmup:
            e2e_ms += dt * 1e3
        a_loc[...] = a_nloc                       # next tick's input = this tick's output (the co-simulation loop)
        a_vel[...] = a_nvel
    t = torch.tensor([e2e_ms], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms_per_step = t.item() / e2e_steps
    e2e_value = n * (n - 1) / (e2e_ms_per_step * 1e-3)
    e.check_peers()                               # a timed-out flag barrier would have produced garbage: fail loudly

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        sm_max_hz = float(peaks.get('sm_max_mhz', 1965.0)) * 1e6
        fp32_peak = SMS * LANES * sm_max_hz / 1e12                      # T FP32-pipe lane-instructions / s
        k1_ms_in_step = ms_pairs / max(stats['pair_launches'], 1)       # overlapped with the cell-list kernels
        k1_evals = stats['pair_evaluations'] / max(stats['pair_launches'], 1)
        k1_ms = k1_ms_in_step
        if world == 1:
            # the pair kernel timed alone (CUDA events on the launch stream): inside the step it shares the SMs with
            # the concurrently running FP64 cell-list kernels, which inflates its span
            ctx.reset_stats()
            ctx.set_profiling(True)
            scratch = np.empty((rows, 3))
            for _ in range(5):
                with torch.cuda.stream(stream):
                    flush.fill_(1)
                ctx.force(1, scratch)
            iso = ctx.stats()
            ctx.set_profiling(False)
            k1_ms = iso['ms_pairs'] / max(iso['pair_launches'], 1)
            k1_evals = iso['pair_evaluations'] / max(iso['pair_launches'], 1)
        # pair terms the launch actually evaluates (the symmetric kernel visits every unordered pair once and applies it
        # to both rows; padded slots included) -- this is what occupies the pipes -- and the ordered pairs it covers
        flat = bool(np.all(w.loc[:, 2] == w.loc[0, 2]) and not np.any(w.vel[:, 2])) and not cfg.get('use_ped_radius', False)
        instr = K1_INSTR_PER_PAIR_2D if flat else K1_INSTR_PER_PAIR
        k1_rate = k1_evals / (k1_ms * 1e-3)
        k1_ordered_rate = (e.hi - e.lo) * (n - 1) / (k1_ms * 1e-3)
        achieved = k1_rate * instr / 1e12
        traffic = None
        prof = os.path.join(ROOT, 'profiles', 'k1_ncu_summary.json')          # ncu --set full capture of k1_sym_pairs
        if os.path.exists(prof):
            with open(prof) as f:
                traffic = json.load(f).get('dram_bytes_per_launch')
        k3_ms = ms_integrate / max(stats['steps'], 1)
        k3_gbs = rows * K3_BYTES_PER_AGENT / (k3_ms * 1e-3) / 1e9 if k3_ms > 0 else None
        line = {
            'metric': 'pair_interactions_per_s', 'value': value, 'unit': 'pair-interactions/s',
            'agent_steps_per_s': value / (n - 1), 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32 pair forces / f64 state, cell-list forces and integration', 'data': 'synthetic',
            'config': {'workload': name, 'n_pedestrians': n, 'rows_per_gpu': rows, 'step_length': w.step_length,
                       'forces': 'all five on', 'l2': 'flushed between timed steps (256 MiB fill, outside the timed spans)',
                       'partition': (f'row blocks over {world} rank(s); per tick an integer reduce-scatter of the pair-force '
                                     f'accumulators and an all-gather of the staged rows (32 B/pedestrian each), transport: '
                                     + ('none (1 rank)' if world == 1 else
                                        ('peer memory over NVLink, fused into k1_sym_finish / k3_integrate (K7)'
                                         if e.peer else 'NCCL reduce_scatter + all_gather'))),
                       'vehicles': ('device-resident: centres advanced and ellipse rings regenerated on the device every tick'
                                    if e.device_vehicles else 'none' if w.veh_center is None else 'host rings')},
            'e2e': {'value': e2e_value, 'unit': 'pair-interactions/s', 'ms_per_step': e2e_ms_per_step,
                    'h2d_bytes_per_step': int(rows * 48), 'd2h_bytes_per_step': int(rows * 48), 'steps': e2e_steps,
                    'api': 'sfm_tick_host (pinned host loc/vel in, new loc/vel out)'},
            'gpu_launches': int(stats['launches']),
            'kernel_ms_per_step': {'pairs_k1': ms_pairs / args.steps, 'segments_cells_k2': ms_sets / args.steps,
                                   'integrate_k3': ms_integrate / args.steps},
            'roofline': {'bound': 'fp32_issue', 'achieved': achieved, 'peak': fp32_peak,
                         'unit': 'T FP32-pipe instr/s', 'frac': achieved / fp32_peak, 'traffic': traffic,
                         'kernel': 'k1_sym_pairs (+ accumulator zero/finish)', 'ms_per_launch': k1_ms,
                         'pair_terms_evaluated_per_s': k1_rate, 'ordered_pairs_covered_per_s': k1_ordered_rate,
                         'frac_ordered_equivalent': k1_ordered_rate * instr / 1e12 / fp32_peak,
                         'ms_per_launch_inside_step': k1_ms_in_step, 'instr_per_pair': instr,
                         'note': 'f_ji = -f_ij exactly, so each unordered pair is evaluated once: achieved/frac count '
                                 'the evaluated pair terms (hardware utilisation); the ordered-pair figures are the '
                                 'useful work the metric counts',
                         'algorithmic': (f'{instr} FP32-pipe instr per pair term (SURVEY 8d: 58 = 3-D radius-on, 81 FLOP, '
                                         '5 MUFU; 48 = the 2-D radius-off specialisation a flat crowd takes)'),
                         'peak_source': f'148 SM x 128 lanes x sm_max_mhz ({peak_kind} MEASURED_PEAKS.json clock); '
                                        'FFMA microbenchmark on this pool: 33.2 T/s (profiles/microbench)'},
            'roofline_hbm': {'bound': 'hbm', 'kernel': 'k3_integrate', 'achieved': k3_gbs,
                             'peak': float(peaks.get('hbm_gbs', 6650.0)), 'unit': 'GB/s',
                             'frac': (k3_gbs / float(peaks.get('hbm_gbs', 6650.0))) if k3_gbs else None,
                             'ms_per_launch': k3_ms, 'peak_kind': peak_kind},
            'clocks': clocks, 'wall_s_timed_loop': wall,
            'ms_per_step_rank0': {'median': float(np.median(per_step)), 'min': float(min(per_step)), 'max': float(max(per_step))},
        }
        if clocks and clocks.get('sm_mhz'):
            line['roofline']['frac_at_sampled_clock'] = achieved / (SMS * LANES * clocks['sm_mhz'] * 1e6 / 1e12)
        if world == 1 and not args.no_cpu_baseline:
            from oracle import cpu_baseline
            cores = os.cpu_count() or 1
            r = cpu_baseline.time_sample(w, cfg, rows_per_core=args.cpu_rows_per_core, cores=cores)
            line['cpu_baseline'] = {
                'value': r['pairs_per_s'], 'unit': 'pair-interactions/s', 'cores': r['cores'], 'kind': 'port',
                'agent_steps_per_s': r['agent_steps_per_s'], 'seconds': r['seconds'],
                'sample': f"one full tick for {r['rows']} of {n} rows (each against all {n} pedestrians + the border/"
                          f"obstacle sets), float64 numpy oracle, {r['cores']} forked workers"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cfg3', choices=['cfg3', 'cfg4', 'cfg5'])
    ap.add_argument('--peds', dest='n', type=int, default=None, help='override the pedestrian count')
    ap.add_argument('--cpu-rows-per-core', type=int, default=384,
                    help='rows of the CPU sample per host core (cpu_baseline leg: ~10 s; the reference arm uses a quarter per step)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
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
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    if args.impl == 'reference':
        run_reference(args, args.gpus, rank)
    else:
        run_ours(args, world, rank, local_rank)


if __name__ == '__main__':
    main()
