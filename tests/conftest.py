import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')




@pytest.fixture(scope='session')
def cstages():
    """oracle/c_stages.c compiled for the host (the C restatement of the front-end arithmetic)."""
    from tests import util
    return util.load_cstages()


@pytest.fixture(scope='session')
def emul():
    """tests/host_emul: the YSMR_HD logic of the CUDA sources compiled for the host CPU."""
    from tests import util
    return util.load_emul()


@pytest.fixture(scope='session')
def golden_dir():
    return os.path.join(ROOT, 'tests', 'golden')
