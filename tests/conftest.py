import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def scenes():
    from pt_three_ways_b200 import scenefile
    names = ["cornell", "suzanne", "ce", "single-sphere", "multi-sphere", "example1", "bbc-owl"]
    return {n: scenefile.load(os.path.join(GOLDEN, "scenes", n + ".ptscene")) for n in names}


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_binding as ob
    ob.lib()
    return ob


@pytest.fixture(scope="session")
def capi():
    from pt_three_ways_b200 import capi as module
    module.lib()
    return module
