"""Row partitioning across ranks (world_size 2, gloo, CPU): the host-side logic of the multi-GPU path.

The device kernels cannot run here, so each rank evaluates its row block with the oracle as the stand-in checker; the
test verifies what the engine's plumbing must guarantee: blocks tile [0, N) without gaps, the all-gathered staging
layout [world][planes][rows_pad] reproduces every rank's rows, and concatenating the per-rank forces equals the
single-process result.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sfm_b200 import engine


def test_partition_bounds():
    for n, world in ((10, 3), (65536, 8), (5, 8), (262144, 4), (1000003, 8)):
        b = engine.partition_rows(n, world)
        assert b[0] == 0 and b[-1] == n and (np.diff(b) >= 0).all()
        tiles = -(-n // 256)
        if tiles >= world:           # whole tiles per rank: block sizes differ by at most one tile (+ the ragged tail)
            assert (b[:-1] % 256 == 0).all() and np.diff(b).max() - np.diff(b).min() <= 256 + 255
        else:
            assert np.diff(b).max() - np.diff(b).min() <= 1
        pad = engine.padded_rows(b)
        assert pad % 256 == 0 and pad >= np.diff(b).max() and pad - np.diff(b).max() < 256


def _rank_main(rank, world, port, n, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, 'carla-social-force-model_b200')]
    from oracle import sfm_oracle as O
    from sfm_b200 import synth
    w = synth.make_config(2, n=n)
    bounds = engine.partition_rows(w.n, world)
    rows_pad = engine.padded_rows(bounds)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    # this rank's staged block: 15 float32 planes x rows_pad -- position hi parts (2^-6 m lattice), lo parts, radius,
    # lambda * velocity, non-planar flag, positions relative to the origin of the row's 64-row run, run origins + compact
    # flags in the first 16 slots of every 256-row tile (csrc/sfm_common.cuh) -- restated on the host
    PX, PXL, PR, PVX, PFLAG, PXR, PMETA, NPLANES = 0, 3, 6, 7, 10, 11, 14, 15
    block = torch.zeros(NPLANES, rows_pad, dtype=torch.float32)
    hi_part = np.rint(w.loc[lo:hi] * 64.0) / 64.0
    block[PX:PX + 3, :hi - lo] = torch.from_numpy(hi_part.T.astype(np.float32))
    block[PXL:PXL + 3, :hi - lo] = torch.from_numpy((w.loc[lo:hi] - hi_part.astype(np.float32)).T.astype(np.float32))
    block[PR, :hi - lo] = torch.from_numpy(w.radius[lo:hi].astype(np.float32))
    block[PVX:PVX + 3, :hi - lo] = torch.from_numpy((2.0 * w.vel[lo:hi]).T.astype(np.float32))
    block[PX:PX + 2, hi - lo:] = 1.0e15                             # pad rows sit far away
    block[PXR:PXR + 2, hi - lo:] = 1.0e15
    for run in range(rows_pad // 64):
        rows = w.loc[lo + 64 * run:min(hi, lo + 64 * (run + 1))]
        meta = 256 * (run // 4) + 4 * (run % 4)
        block[PMETA, meta + 3] = 1.0                                # a run without live rows counts as compact
        if len(rows) == 0:
            continue
        c = np.rint(0.5 * (rows.min(axis=0) + rows.max(axis=0)) * 64.0) / 64.0
        block[PXR:PXR + 3, 64 * run:64 * run + len(rows)] = torch.from_numpy((rows - c).T.astype(np.float32))
        block[PMETA, meta:meta + 3] = torch.from_numpy(c.astype(np.float32))
        block[PMETA, meta + 3] = float(np.abs(rows - c).max() <= 16.0)
    gathered = torch.zeros(world, NPLANES, rows_pad, dtype=torch.float32)
    dist.all_gather_into_tensor(gathered.view(-1), block.view(-1))
    # rebuild the global crowd from the gathered layout (hi + lo is exact) and compute this rank's rows
    rows_of = lambda q: int(bounds[q + 1] - bounds[q])                                        # noqa: E731
    loc = np.concatenate([(gathered[q, PX:PX + 3, :rows_of(q)].numpy().astype(np.float64)
                           + gathered[q, PXL:PXL + 3, :rows_of(q)].numpy().astype(np.float64)).T for q in range(world)])
    vel = np.concatenate([gathered[q, PVX:PVX + 3, :rows_of(q)].numpy().T for q in range(world)]).astype(np.float64) / 2.0
    rad = np.concatenate([gathered[q, PR, :rows_of(q)].numpy() for q in range(world)]).astype(np.float64)
    assert np.array_equal(loc, w.loc) and np.array_equal(vel, w.vel) and np.array_equal(rad, w.radius)
    # the run-local copies say the same positions to float32 rounding at the run's extent (origin + xr)
    for q in range(world):
        for run in range((rows_of(q) + 63) // 64):
            meta = 256 * (run // 4) + 4 * (run % 4)
            c = gathered[q, PMETA, meta:meta + 3].numpy().astype(np.float64)
            k = min(64, rows_of(q) - 64 * run)
            xr = gathered[q, PXR:PXR + 3, 64 * run:64 * run + k].numpy().astype(np.float64).T
            want = w.loc[int(bounds[q]) + 64 * run:int(bounds[q]) + 64 * run + k]
            bound = 2.0 ** -24 * max(1.0, np.abs(want - c).max())
            assert np.abs(c + xr - want).max() <= bound
    f = O.pedestrian_force(loc, vel, rad, rows=np.arange(lo, hi))
    np.save(os.path.join(out_dir, f'f{rank}.npy'), f)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_row_partition_matches_single_process(tmp_path):
    from oracle import sfm_oracle as O
    from sfm_b200 import synth
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    n = 300                                                          # ragged: 150 + 150 rows in 256-row blocks
    mp.spawn(_rank_main, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    w = synth.make_config(2, n=n)
    whole = O.pedestrian_force(w.loc, w.vel, w.radius)
    parts = np.concatenate([np.load(tmp_path / f'f{r}.npy') for r in range(2)])
    np.testing.assert_allclose(parts, whole, rtol=1e-12, atol=1e-13)     # chunk shapes differ, sums reassociate
