"""The accuracy argument of the pair kernel's run-local path, pinned on the CPU (profiles/local_origin_emulation.py).

The kernel forms d = xr_j - m_i (one float32 per coordinate relative to the origin of the partner's 64-slot run) only for
tile pairs whose bounding boxes are far apart, and the double-single form everywhere else (csrc/sfm_common.cuh).  The
emulation forms d exactly as the kernel does -- numpy float32, same operation order, same tile rule -- and evaluates the
rest of forces.py:74-117 in float64 with the oracle, so what it measures is the staging error alone, against the
1e-4 rel / 1e-5 abs force tolerance.
"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'profiles'))


def test_emulation_uses_the_kernel_constants():
    import local_origin_emulation as E
    text = open(os.path.join(ROOT, 'carla-social-force-model_b200', 'csrc', 'sfm_common.cuh')).read()
    const = lambda name: float(re.search(rf'constexpr \w+ {name} = ([0-9.]+)f?;', text).group(1))      # noqa: E731
    assert (E.LOCAL_SEP, E.LOCAL_SEP_FACTOR, E.LOCAL_LIMIT) == (const('LOCAL_SEP'), const('LOCAL_SEP_FACTOR'), const('LOCAL_LIMIT'))
    assert const('SUB_ROWS') == 64 and const('POS_LATTICE') == 64.0


def test_staging_error_of_the_local_path_stays_at_the_double_single_level():
    """cfg2 (N = 4,096 on 64 x 64 m, shifted 500 m from the origin, off the float32 lattice), rows incl. the ones with the
    closest neighbours: a good share of the pairs takes the local path and the staging error stays far below the tolerance
    -- because every pair that is close enough to matter is still evaluated in double-single form."""
    import local_origin_emulation as E
    out = E.emulate(cfg=2, n=4096, n_rows=48, verbose=False)
    assert out['local_share'] > 0.3, out
    assert out['double-single'] < 0.05 and out['local'] < 0.05, out


def test_hilbert_order_keeps_runs_compact():
    """The host restatement of the staged slot order (csrc/k8_order.cuh): a stable permutation under which every run of 64
    consecutive rows of a 1 pedestrian / m^2 crowd spans a few metres -- including the rows on the far edges of the
    bounding square (a curve over a square larger than the crowd would send them to the end of the order) -- while in
    the crowd's own (random) row order a run spans the whole square."""
    import numpy as np
    import local_origin_emulation as E
    from sfm_b200 import synth
    w = synth.make_config(2)                                     # N = 4,096 on 64 x 64 m
    order = E.hilbert_order(w.loc)
    assert sorted(order.tolist()) == list(range(w.n))
    runs = w.loc[order, :2].reshape(-1, 64, 2)
    extent = (runs.max(axis=1) - runs.min(axis=1)).max(axis=1)
    assert extent.max() < 24.0 and np.median(extent) < 12.0, (extent.max(), np.median(extent))
    unordered = w.loc[:, :2].reshape(-1, 64, 2)
    assert (unordered.max(axis=1) - unordered.min(axis=1)).max(axis=1).min() > 40.0
    # stable: rows in the same cell keep their order
    same = np.zeros((8, 3))
    assert E.hilbert_order(same).tolist() == list(range(8))
