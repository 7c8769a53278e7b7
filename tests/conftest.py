import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)
sys.dont_write_bytecode = True


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def state_dicts():
    """One deterministic state_dict per preset (oracle/weights.py, seed 5), built lazily."""
    import weights
    cache = {}

    def get(preset):
        if preset not in cache:
            cache[preset] = weights.make_state_dict(preset, seed=5)
        return cache[preset]
    return get
