import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_mod():
    """The CPU oracle (test infrastructure); built on first use."""
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def cfg_codes():
    from bp_osd_b200 import codes
    cache = {}

    def get(cfg):
        if cfg not in cache:
            cache[cfg] = codes.config_code(cfg)
        return cache[cfg]
    return get


@pytest.fixture(scope="session")
def cuda_lib():
    """Build (if needed) and load the CUDA library; a missing library is an error, never a skip."""
    from bp_osd_b200 import build, _capi
    build.build()
    return _capi.load()
