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
        self.params_key = None
        self.owner = {native.BORDER: None, native.STATIC_OBSTACLE: None, native.DYNAMIC_OBSTACLE: None}
        self.set_version = {native.BORDER: None, native.STATIC_OBSTACLE: None, native.DYNAMIC_OBSTACLE: None}

    def bind_params(self, sfm_config, step_length):
        params = native.params_from_config(sfm_config, step_length,
                                           enable={name: True for name in native.FORCE_CLASSES})
        key = bytes(params)
        if key != self.params_key:
            # a changed perception threshold invalidates resident obstacle sets (their cutoff is baked in)
            self.ctx.set_params(params)
            self.params_key = key
            self.owner = {k: None for k in self.owner}
        return params

    def upload_peds(self, peds, mode_codes=None):
        self.ctx.upload_state(*peds.device_columns(mode_codes))

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
        _session.ctx.close()
    _session = None
