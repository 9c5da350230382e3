"""cfg5 trajectory agreement (BASELINE.json configs[4]): N = 1,048,576 pedestrians.

  torchrun --nproc-per-node 8 profiles/cfg5_agreement.py --phase multi  --steps K   # 8 ranks, state dumped per rank
  python profiles/cfg5_agreement.py --phase single --steps K                          # 1 GPU, same K ticks + oracle sample
  python profiles/cfg5_agreement.py --phase compare                                   # bitwise comparison, JSON summary

The integer (fixed-point) pair-force accumulation makes the result independent of how the tiles are spread over ranks, so
the 8-GPU and the 1-GPU trajectories must agree BIT FOR BIT; the oracle check compares the float32 pair force of a row
sample with the float64 restatement of forces.py:74-117 at the BASELINE tolerance (1e-4 relative + 1e-5 absolute).
"""
import argparse
import json
import os
import sys
import time
import tomllib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'carla-social-force-model_b200')]
OUT = os.path.join(ROOT, 'gpurun_out', 'cfg5_agreement')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--phase', required=True, choices=['multi', 'single', 'compare'])
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--peds', dest='n', type=int, default=1048576)
    ap.add_argument('--oracle-rows', type=int, default=48)
    ap.add_argument('--stride', type=int, default=16, help='rows kept for the comparison: every stride-th global row')
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    if args.phase == 'compare':
        single = np.load(os.path.join(OUT, 'single.npz'))
        world = int(np.load(os.path.join(OUT, 'rank0.npz'))['world'])
        parts = [np.load(os.path.join(OUT, f'rank{r}.npz')) for r in range(world)]
        loc = np.concatenate([p['loc'] for p in parts])
        vel = np.concatenate([p['vel'] for p in parts])
        res = {'rows_compared': int(len(loc)), 'n': int(single['n']), 'world': world, 'steps': int(single['steps']),
               'bitwise_equal_loc': bool(np.array_equal(loc, single['loc'])),
               'bitwise_equal_vel': bool(np.array_equal(vel, single['vel'])),
               'max_abs_dloc': float(np.abs(loc - single['loc']).max()), 'max_abs_dvel': float(np.abs(vel - single['vel']).max()),
               'ms_per_step_multi': float(parts[0]['ms_per_step']), 'ms_per_step_single': float(single['ms_per_step']),
               'oracle_rows': int(single['oracle_rows']), 'oracle_worst_err_over_tol': float(single['oracle_worst']),
               'oracle_max_abs_err': float(single['oracle_max_abs'])}
        print(json.dumps(res))
        return

    import torch
    from sfm_b200 import engine as eng, native, synth
    with open(os.path.join(ROOT, 'carla-social-force-model_b200', 'config', 'sfm_config.toml'), 'rb') as f:
        cfg = tomllib.load(f)
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if args.phase == 'multi':
        torch.distributed.init_process_group('nccl', device_id=torch.device('cuda', local))
    w = synth.make_config(5, n=args.n)
    # rows staged in ROW order on both sides: then the tile pairs of the 8-rank run are those of the single-GPU run and the
    # trajectories agree bit for bit.  (With the rank-local Hilbert staging -- Engine's default -- a rank's tiles differ from
    # the single-GPU run's, and the agreement is to float32 rounding of the tile partials; bench.py's parity block replays a
    # tick under the ranks' own orders for the bitwise check in that mode.)
    e = eng.Engine(cfg, w.step_length, device=local, reorder_every=int(os.environ.get('SFM_REORDER_EVERY', '0')))
    e.load(w)
    extra = {}
    if args.phase == 'single':
        # float32 pair force of the initial state against the float64 oracle on a row sample
        from oracle import sfm_oracle as O
        rows = np.linspace(0, w.n - 1, args.oracle_rows).astype(np.int64)
        f_dev = e.ctx.force(native.PEDESTRIAN)[rows]
        pp = O.moussaid_params(cfg['pedestrian_force'], O.PED_DEFAULTS)
        want, risk = O.pedestrian_force(w.loc, w.vel, w.radius, pp, cfg.get('use_ped_radius', False), rows=rows, chunk=4,
                                        return_risk=True)
        err = np.abs(f_dev - want)
        tol = 1e-5 + 1e-4 * np.abs(want) + risk[:, None]
        extra = dict(oracle_rows=len(rows), oracle_worst=float((err / tol).max()), oracle_max_abs=float(err.max()))
    e.step(2, True)                                                    # warm-up ticks are part of the trajectory
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e.step(args.steps - 2, True)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / max(args.steps - 2, 1)
    loc, vel = e.local_state()
    keep = (np.arange(e.lo, e.hi) % args.stride) == 0          # a fixed global row sample (files travel back from the box)
    loc, vel = loc[keep], vel[keep]
    name = f'rank{rank}.npz' if args.phase == 'multi' else 'single.npz'
    np.savez(os.path.join(OUT, name), loc=loc, vel=vel, world=world, steps=args.steps, ms_per_step=ms, n=w.n, **extra)
    if args.phase == 'multi':
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    if rank == 0:
        print(f'{args.phase}: n={w.n} world={world} steps={args.steps} {ms:.2f} ms/step {extra}', flush=True)


if __name__ == '__main__':
    main()
