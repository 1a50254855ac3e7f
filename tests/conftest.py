import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.dirname(os.path.abspath(__file__)) not in sys.path:
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))     # tests/tree_helpers.py


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `-m gpu` under gpurun)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def ref_runs(golden_dir):
    import json
    with open(os.path.join(golden_dir, "ref_runs.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def post_golden(golden_dir):
    import numpy as np
    return np.load(os.path.join(golden_dir, "post.npz"))
