"""The C-ABI library loads without a GPU and exports exactly what include/sfm_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from sfm_b200 import native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def built():
    return native.build()


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'sfm_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return set(re.findall(r'\b(sfm_[a-z_0-9]+)\s*\(', text))


def test_header_matches_binding_table():
    assert declared_symbols() == set(native.SYMBOLS)


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(built)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.sfm_abi_version() == native.ABI_VERSION


def test_struct_layout_matches_header():
    # sfm_params: 5 doubles + 3 x 7 doubles + 6 int32 ; sfm_stats: 12 x 8 bytes
    assert ctypes.sizeof(native.MoussaidParams) == 56
    assert ctypes.sizeof(native.Params) == 5 * 8 + 3 * 56 + 6 * 4
    assert ctypes.sizeof(native.Stats) == 96


def test_fails_loudly_without_device(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    with pytest.raises(native.SfmError):
        native.Context(0)
    from sfm_b200 import engine
    with pytest.raises(native.SfmError):
        engine.Engine({}, 0.05)


def test_params_follow_reference_key_quirks(sfm_config):
    p = native.params_from_config(sfm_config, 0.05)
    assert p.max_speed_factor == 1.3 and p.tau == 0.5                 # shipped keys are NOT the ones read
    q = native.params_from_config(dict(sfm_config, max_speed_multiplier=9.0, max_speed_factor=1.5,
                                       goal_force={'tau': 0.25}), 0.05)
    assert q.max_speed_factor == 1.5 and q.tau == 0.25
    assert (p.border_a, p.border_b) == (6.0, 0.3)
    assert native.params_from_config({}, 0.05).border_a == 3.0         # code defaults differ from shipped values
    assert list(p.enable) == [1, 1, 1, 1, 1]
    assert p.dynamic_obs.perception_threshold == 50 and p.static_obs.lambda_weight == 2.3
    with pytest.raises(KeyError):
        native.params_from_config({}, 0.05, strict=True)
