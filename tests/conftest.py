"""pytest configuration: marker registration and import paths.

The drop-in package directory is put on ``sys.path`` the same way the reference expects its own directory to be
(modules import each other by bare name, forces.py:7-8).
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'carla-social-force-model_b200')
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def sfm_config():
    import tomllib
    with open(os.path.join(PKG, 'config', 'sfm_config.toml'), 'rb') as f:
        return tomllib.load(f)
