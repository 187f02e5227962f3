import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")
    config.addinivalue_line("markers", "multigpu: needs >= 2 GPUs")
    config.addinivalue_line("markers", "gpu_experimental: needs a B200; covers opt-in code that has not yet run on hardware "
                            "(NOT part of `-m gpu`; run with `-m gpu_experimental` under gpurun with a timeout)")


def pytest_collection_modifyitems(config, items):
    """`gpu_experimental` tests run only when asked for by name (`-m gpu_experimental`): neither the CPU suite (`-m "not gpu"`) nor the
    GPU suite (`-m gpu`) may pick them up."""
    if "gpu_experimental" in (config.getoption("-m") or ""):
        return
    skip = pytest.mark.skip(reason="opt-in code not yet run on hardware: select with -m gpu_experimental")
    for item in items:
        if "gpu_experimental" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def b200():
    import elmerfem_b200 as B
    if not os.path.exists(B.LIB_PATH):
        B.build()
    B.lib()
    return B
