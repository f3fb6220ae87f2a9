import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    from helpers import oracle_library
    return oracle_library()


@pytest.fixture(scope="session")
def cuda_lib():
    """The product library.  GPU tests must exercise the CUDA path; there is no fallback."""
    import gigapaxos_b200
    return gigapaxos_b200.load_library()
