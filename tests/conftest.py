import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with gzip.open(os.path.join(GOLDEN, name), "rb") as f:
        return json.loads(f.read().decode())


@pytest.fixture(scope="session")
def golden_windows():
    return load_golden("windows.json.gz")


@pytest.fixture(scope="session")
def golden_spoa():
    return load_golden("spoa_global_consensus.json.gz")


@pytest.fixture(scope="session")
def golden_windows_large():
    return load_golden("windows_large.json.gz")
