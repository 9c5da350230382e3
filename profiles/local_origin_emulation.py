"""TEST INFRASTRUCTURE, not product code (imports the CPU oracle; run by tests/test_staging_emulation.py and by hand).

CPU emulation of the tile-local staging of the pair kernel (round 2, K1s "local" path): what does the float32
rounding of the staged coordinates ALONE cost against the 1e-4 rel / 1e-5 abs force tolerance?

For sampled rows i and every j, the position difference d is formed exactly as the kernel forms it (numpy float32
arithmetic, same operation order), everything after that in float64 with the oracle's own code -- so the reported error
is the staging error in isolation, not the kernel's arithmetic error.

  double-single (all tiles)   d = (hi_j - hi_i) + (lo_j - lo_i)
  local (compact 64-row runs) d = xr_j - m_i,   xr_j = float32(x_j - c_run),   m_i = (hi_i - c_run) + lo_i

Usage: python profiles/local_origin_emulation.py [cfg] [n] [rows]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'carla-social-force-model_b200')]
from oracle import sfm_oracle as O            # noqa: E402
from sfm_b200 import synth                    # noqa: E402

f32 = np.float32
LOCAL_SEP, LOCAL_SEP_FACTOR, LOCAL_LIMIT = 1.0, 0.25, 64.0      # csrc/sfm_common.cuh


def hilbert_order(loc, cell=0.25, bits=16):
    """Host restatement of the staged slot order the device builds (csrc/k8_order.cuh): rows along a Hilbert curve over
    their xy positions, stable."""
    xy = np.asarray(loc, dtype=np.float64)[:, :2]
    if len(xy) == 0:
        return np.zeros(0, dtype=np.int64)
    lo = xy.min(axis=0)
    span = float((xy.max(axis=0) - lo).max())
    # the curve's square is the crowd's bounding square exactly, cut into the smallest power-of-two number of cells that
    # are no larger than `cell` (a square larger than the crowd would put the rows on its far edges into another
    # top-level quadrant of the curve, i.e. at the end of the order)
    n = 2
    while n * float(cell) < span and n < (1 << bits):
        n <<= 1
    q = np.minimum(((xy - lo) * (n / span if span > 0 else 0.0)).astype(np.int64), n - 1)
    x, y = q[:, 0].copy(), q[:, 1].copy()
    d = np.zeros(len(xy), dtype=np.int64)
    s = n >> 1
    while s > 0:
        rx = (x & s) > 0
        ry = (y & s) > 0
        d += s * s * ((3 * rx.astype(np.int64)) ^ ry.astype(np.int64))
        flip = ~ry & rx
        x = np.where(flip, n - 1 - x, x)
        y = np.where(flip, n - 1 - y, y)
        x, y = np.where(~ry, y, x), np.where(~ry, x, y)
        s >>= 1
    return np.argsort(d, kind='stable')


def staged(loc, origin):
    rel = loc - origin
    hi = (np.rint(rel * 64.0) / 64.0).astype(f32)
    lo = (rel - hi.astype(np.float64)).astype(f32)
    return hi, lo


def tile_origins(loc, origin, tile=64):
    """Per 64-row run: bounding-box centre on the 2^-6 m lattice (relative to the origin), xr = float32(x - origin - c)."""
    n = len(loc)
    rel = loc - origin
    c = np.zeros((n, 3))
    for t in range(0, n, tile):
        blk = rel[t:t + tile]
        c[t:t + tile] = np.rint(0.5 * (blk.min(0) + blk.max(0)) * 64.0) / 64.0
    xr = (rel - c).astype(f32)
    return c.astype(f32), xr


def force_from_d(d, vel_i, vel, own, p):
    d = d.astype(np.float64)
    direction, length = O.normalize(d)
    force, *_ = O._moussaid(direction, length, vel_i[:, None, :] - vel[None, :, :], p)
    force = np.where(own[..., None], 0.0, force)
    return force.sum(1)


def emulate(cfg=3, n=None, n_rows=96, verbose=True):
    """Returns {'local_share': ..., 'double-single': worst error / tolerance, 'local': worst error / tolerance}."""
    say = print if verbose else (lambda *a, **k: None)
    w = synth.make_config(cfg, n=n, scale_sets=False) if n else synth.make_config(cfg)
    rng = np.random.default_rng(7)
    shift = np.array([512.1, -498.9, 0.0])
    loc = w.loc + shift + np.concatenate((rng.uniform(-0.03, 0.03, (w.n, 2)), np.zeros((w.n, 1))), axis=1)
    vel = w.vel
    order = hilbert_order(loc)
    loc, vel = loc[order], vel[order]
    origin = np.rint(0.5 * (loc.min(0) + loc.max(0)))
    hi, lo = staged(loc, origin)
    c, xr = tile_origins(loc, origin)
    ext = np.abs(xr[:, :2]).reshape(-1, 64, 2).max(1)
    say(f'cfg{cfg} N={w.n}: run half-extent median {np.median(ext):.1f} m, 99 % {np.quantile(ext, 0.99):.1f} m, '
          f'max {ext.max():.1f} m')
    far = np.argsort(-np.abs(loc[:, :2] - origin[:2]).max(1))[:n_rows // 3]
    from scipy.spatial import cKDTree
    nn = cKDTree(loc[:, :2]).query(loc[:, :2], k=2)[0][:, 1]
    close = np.argsort(nn)[:n_rows // 3]                         # the rows with the closest neighbours (down to millimetres)
    rows = np.unique(np.concatenate((rng.choice(w.n, n_rows - len(far) - len(close), replace=False), far, close)))
    say(f'  sampled rows: {len(rows)} (a third each: uniform, farthest from the origin, closest neighbour -- from '
          f'{nn[close].min() * 1e3:.1f} mm)')
    p = dict(O.PED_DEFAULTS)
    exact, risk = O.pedestrian_force(loc, vel, w.radius, rows=rows, return_risk=True)
    tol = 1e-5 + 1e-4 * np.abs(exact) + risk[:, None]
    worst = {'double-single': 0.0, 'local': 0.0}
    ratios = {'double-single': [], 'local': []}
    # which (row, partner) pairs take the local path: the bounding boxes of the two 256-row tiles are at least
    # max(LOCAL_SEP, LOCAL_SEP_FACTOR * widest run of the partner tile) apart (k1_sym.cuh); everything else -- the row's own
    # tile, adjacent tiles, tiles with a spread-out run -- takes the double-single path
    rel = (loc - origin)[:, :2]
    tiles = rel.reshape(-1, 256, 2)
    t_lo, t_hi = tiles.min(1), tiles.max(1)
    run_ext = np.where(ext.max(1) <= LOCAL_LIMIT, ext.max(1), np.inf).reshape(-1, 4).max(1)     # widest run of each tile
    taken = 0
    for s in range(0, len(rows), 32):
        r = rows[s:s + 32]
        own = r[:, None] == np.arange(w.n)[None, :]
        ti = r // 256
        gap = np.maximum(t_lo[None, :, :] - t_hi[ti, None, :], t_lo[ti, None, :] - t_hi[None, :, :]).max(-1)   # (rows, tiles)
        local_tile = (gap >= np.maximum(LOCAL_SEP, LOCAL_SEP_FACTOR * run_ext)[None, :]) & \
                     (ti[:, None] != np.arange(len(run_ext))[None, :])
        local_pair = np.repeat(local_tile, 256, axis=1)
        taken += int(local_pair.sum())
        d_ds = (hi[None, :, :] - hi[r, None, :]) + (lo[None, :, :] - lo[r, None, :])          # float32 throughout
        m = (hi[r, None, :] - c[None, :, :]) + lo[r, None, :]                                  # (rows, N, 3) float32
        d_loc = np.where(local_pair[..., None], xr[None, :, :] - m, d_ds)
        for name, d in (('double-single', d_ds), ('local', d_loc)):
            ratio = np.abs(force_from_d(d, vel[r], vel, own, p) - exact[s:s + 32]) / tol[s:s + 32]
            worst[name] = max(worst[name], float(ratio.max()))
            ratios[name].append(ratio.max(1))
    say(f'  pairs on the local path: {taken / (len(rows) * (w.n - 1)):.3f} of all pairs of the sampled rows')
    for name, v in worst.items():
        q = np.concatenate(ratios[name])
        say(f'  {name:14s} staging error / tolerance over {len(rows)} rows: median {np.median(q):.4f}, 99 % '
              f'{np.quantile(q, 0.99):.4f}, worst {v:.4f}')
    return {'local_share': taken / (len(rows) * (w.n - 1)), 'double-single': worst['double-single'], 'local': worst['local']}


def main():
    cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    n = int(sys.argv[2]) if len(sys.argv) > 2 else None
    n_rows = int(sys.argv[3]) if len(sys.argv) > 3 else 96
    emulate(cfg, n, n_rows)


if __name__ == '__main__':
    main()
