"""Host-side state helpers with the call surface of the reference's ``stateutils.py`` (numpy in, numpy out).

These are the small O(N) utilities callers of the Social Force Model use around the hot path (spawn rotation,
gap-acceptance check, arrival test).  The O(N^2) members (``all_diffs`` / ``all_sums``) are kept for API completeness
only: the device path never materialises pair matrices (see csrc/k1_ped_pairs.cuh).
Reference semantics cited per function (stateutils.py line numbers of felixlutz/carla-social-force-model).
"""
from typing import Tuple

import numpy as np


def normalize(array, axis=-1) -> Tuple[np.ndarray, np.ndarray]:
    """Unit vectors and their original lengths; zero vectors stay zero and report length 0 (stateutils.py:78-92)."""
    lengths = np.linalg.norm(array, axis=axis)
    divisor = np.where(lengths == 0.0, 1.0, lengths)
    return array / np.expand_dims(divisor, axis), lengths


def desired_directions(state):
    """Unit xy direction from ``loc`` to ``next_waypoint`` with a zero z column (stateutils.py:7-15)."""
    heading, _ = normalize(state['next_waypoint'][:, :2] - state['loc'][:, :2])
    out = np.zeros((len(heading), 3))
    out[:, :2] = heading
    return out


def cap_velocity(desired_velocity, max_velocity):
    """Scale each velocity down to at most ``max_velocity`` (3-D norm, zero speed treated as 1; stateutils.py:18-23)."""
    speed = np.linalg.norm(desired_velocity, axis=-1)
    speed = np.where(speed == 0.0, 1.0, speed)
    scale = np.minimum(1.0, max_velocity / speed)
    return desired_velocity * scale[..., np.newaxis]


def speeds(state):
    """3-D speed of every pedestrian (stateutils.py:26-29)."""
    return np.linalg.norm(state['vel'], axis=1)


def _off_diagonal_columns(n):
    """Column index j of the k-th kept entry of row i once the diagonal is dropped: j = k + (k >= i)."""
    k = np.arange(n - 1)[np.newaxis, :]
    i = np.arange(n)[:, np.newaxis]
    return k + (k >= i)


def all_diffs(array, remove_diagonal=True, keep_dims=True):
    """``D[i, j] = a[j] - a[i]``; with the diagonal removed each row keeps j ascending, skipping i (stateutils.py:32-53)."""
    array = np.asarray(array)
    n = array.shape[0]
    full = array[np.newaxis, ...] - array[:, np.newaxis, ...]
    if not remove_diagonal:
        return full
    cols = _off_diagonal_columns(n)
    kept = full[np.arange(n)[:, np.newaxis], cols]
    if keep_dims:
        return kept
    return kept.reshape((n * (n - 1),) + array.shape[1:])


def all_sums(array, remove_diagonal=True, keep_dims=True):
    """``S[i, j] = a[i] + a[j]`` with the same diagonal handling as ``all_diffs`` (stateutils.py:56-75)."""
    array = np.asarray(array)
    n = array.shape[0]
    full = array[:, np.newaxis, ...] + array[np.newaxis, ...]
    if not remove_diagonal:
        return full
    return full[np.arange(n)[:, np.newaxis], _off_diagonal_columns(n)]


def angle_diff_2d(vecs1, vecs2):
    """``atan2(v1) - atan2(v2)`` on the xy components with one +-2 pi correction into [-pi, pi] (stateutils.py:95-128)."""
    vecs1, vecs2 = np.asarray(vecs1), np.asarray(vecs2)
    diff = np.arctan2(vecs1[..., 1], vecs1[..., 0]) - np.arctan2(vecs2[..., 1], vecs2[..., 0])
    two_pi = 2 * np.pi
    if np.ndim(diff) == 0:
        if diff > np.pi:
            diff -= two_pi
        elif diff < -np.pi:
            diff += two_pi
        return diff
    diff = np.where(diff > np.pi, diff - two_pi, diff)
    return np.where(diff < -np.pi, diff + two_pi, diff)
