import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


GOLDEN_CASES = ["ont_mem", "ont_bal", "clr_ratio", "hifi"]


@pytest.fixture(scope="session")
def golden():
    """Loader for tests/golden/<case>: regenerated input + the reference's per-stage dumps (cached)."""
    import golden_io
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = golden_io.load_case(name)
        return cache[name]
    return get
