#!/usr/bin/env python
"""A complete headless run on one B200: the reference's scenario loop without CARLA, everything device-resident.

    python examples/headless_run.py [--peds 4096] [--ticks 200] [--out /tmp/sfm_out]

Builds a synthetic crowd with routes, mode machines and crossing vehicles (sfm_b200.synth), runs
SimulationRunner.tick's sequence on the device (vehicles -> mode machines + gap acceptance -> forces, velocities,
waypoint hand-overs, positions), records every 10th tick and writes the reference's four CSV files.
"""
import argparse
import os
import sys
import time
import tomllib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'carla-social-force-model_b200')]

from sfm_b200 import synth                      # noqa: E402
from sfm_b200.headless import HeadlessRunner    # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--peds', type=int, default=4096)
    ap.add_argument('--ticks', type=int, default=200)
    ap.add_argument('--out', default='/tmp/sfm_out')
    args = ap.parse_args()
    with open(os.path.join(ROOT, 'carla-social-force-model_b200', 'config', 'sfm_config.toml'), 'rb') as f:
        cfg = tomllib.load(f)
    side = float(args.peds) ** 0.5
    w, life = synth.make_lifecycle(n=args.peds, side=side, n_vehicles=max(3, args.peds // 256))
    frames = args.ticks // 10 + 1
    run = HeadlessRunner(cfg, w, life, device_vehicles=True, record_every=10, record_capacity=frames)
    t0 = time.perf_counter()
    run.run(args.ticks)
    run.ctx.synchronize()
    sec = time.perf_counter() - t0
    counters = run.ctx.lifecycle_counters()
    out = run.write_csv(args.out, f'headless-n{args.peds}')
    print(f'{args.ticks} ticks of {args.peds} pedestrians in {sec:.3f} s ({args.peds * args.ticks / sec:.3e} agent-steps/s); '
          f'{counters}; CSV files in {out}')


if __name__ == '__main__':
    main()
