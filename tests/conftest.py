import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture()
def cfg():
    """A pristine hot-path config; overrides are dropped after the test."""
    from eve_b200.config import DefaultConfig
    c = DefaultConfig()
    c.reset()
    yield c
    c.reset()
