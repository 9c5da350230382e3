"""Process-wide device session shared by the ``forces.Force`` objects and ``PedestrianSimulation``.

The reference's Force classes are independent Python objects; here they are thin views onto one device context (the
reference runs one ``PedestrianSimulation`` per process, SURVEY.md section 8b "Threading").  The session remembers which
object's parameters / point sets are currently resident so that composing forces by hand stays correct: every call
re-binds what differs.
"""
from __future__ import annotations

import os

from . import native

_session = None


class Session:
    def __init__(self, device):
        self.ctx = native.Context(device)
        # rows stay in the caller's (spawn) order; below the API they are staged along a Hilbert curve, rebuilt every few
        # ticks, so that the pair kernel's run-local path applies (csrc/k8_order.cuh) -- speed only
        self.ctx.set_reorder_interval(int(os.environ.get('SFM_REORDER_EVERY', '32')))
        self.params_key = None
        self.thresholds = None
        self.owner = {native.BORDER: None, native.STATIC_OBSTACLE: None, native.DYNAMIC_OBSTACLE: None}
        self.set_version = {native.BORDER: None, native.STATIC_OBSTACLE: None, native.DYNAMIC_OBSTACLE: None}
        self.resident_table = None          # ModeTable whose pedestrian rows (and machines) are on the device
        self.machines_version = None        # ... and the table version its device mirror reflects
        self.traffic_version = None
        self.identity_table = None          # ModeTable whose `mode` object pointers the device holds (sfm_tick_records)
        self.pinned = None                  # native.PinnedArray of the PedState table the resident tick copies from

    def pin(self, state):
        """Page-lock the pedestrian table the resident tick hands to ``sfm_tick_records`` (re-done when the table is
        replaced by a spawn / despawn); a failed registration just leaves the copy on the driver's staging path."""
        if self.pinned is not None and self.pinned.array is state:
            return
        if self.pinned is not None:
            self.pinned.release()
        self.pinned = native.PinnedArray(state)

    def set_params(self, params):
        """Make ``params`` the context's parameters unless they already are.  Point sets stay resident across parameter
        changes -- only a changed perception threshold invalidates them (their cutoffs are built from it)."""
        key = bytes(params)
        if key == self.params_key:
            return
        self.ctx.set_params(params)
        self.params_key = key
        thresholds = (params.static_obs.perception_threshold, params.dynamic_obs.perception_threshold)
        if thresholds != self.thresholds:
            self.owner = {k: None for k in self.owner}
            self.thresholds = thresholds

    def bind_params(self, sfm_config, step_length):
        params = native.params_from_config(sfm_config, step_length,
                                           enable={name: True for name in native.FORCE_CLASSES})
        self.set_params(params)
        return params

    def upload_peds(self, peds, mode_codes=None):
        self.ctx.upload_state(*peds.device_columns(mode_codes))
        self.resident_table = None          # a full upload drops the device-side mode machines
        self.identity_table = None

    def bind_set(self, which, owner, version, loader):
        """Make ``owner``'s point set resident for class ``which`` unless it already is (same object, same version)."""
        if self.owner[which] is owner and self.set_version[which] == version:
            return
        loader(self.ctx)
        self.owner[which], self.set_version[which] = owner, version


def get_session():
    """The session on ``SFM_DEVICE`` (default 0).  Raises ``SfmError`` when no B200 is present: no CPU fallback."""
    global _session
    if _session is None:
        _session = Session(int(os.environ.get('SFM_DEVICE', '0')))
    return _session


def reset_session():
    global _session
    if _session is not None:
        if _session.pinned is not None:
            _session.pinned.release()
        _session.ctx.close()
    _session = None
