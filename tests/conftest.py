import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests need libptb200.so and a CUDA device: skip them (rather than fail with
    PTB200_ECUDA) on a box that has neither, so a plain `pytest tests/` is green anywhere."""
    gpu_items = [item for item in items if "gpu" in item.keywords]
    if not gpu_items:
        return
    try:
        from pt_three_ways_b200 import capi
        devices = capi.device_count()
        reason = "no CUDA device visible"
    except Exception as exc:  # library not built
        devices, reason = 0, f"libptb200.so unavailable: {exc}"
    if devices == 0:
        skip = pytest.mark.skip(reason=reason)
        for item in gpu_items:
            item.add_marker(skip)


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
