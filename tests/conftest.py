import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "gpu_pending: needs a CUDA device and has not run on one yet (-m gpu_pending); "
                                       "promoted to `gpu` once validated")
    config.addinivalue_line("markers", "slow: full BASELINE-size cases that take more than a few seconds")


@pytest.fixture(scope="session")
def goldens():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "fish_goldens.json")) as f:
        return json.load(f)
