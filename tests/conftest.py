import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
CONFIGS = os.path.join(GOLDEN, "configs")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def load_cfg(name):
    import json
    with open(os.path.join(CONFIGS, name + ".json"), "r", encoding="utf-8") as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def north_pit():
    return load_cfg("north_pit_mine")


@pytest.fixture(scope="session")
def short_mine():
    return load_cfg("north_pit_mine_short")


@pytest.fixture(scope="session")
def ragged_mine():
    """L=3 != D=2, 1/4/7 shovels and 1/9 dumpers per site, 130 trucks of four types (tests/golden/make_golden_edge.py)."""
    return load_cfg("ragged_mine")


@pytest.fixture(scope="session")
def one_truck_mine():
    return load_cfg("one_truck_mine")


@pytest.fixture(scope="session")
def integer_mine():
    """Every duration a whole number of minutes: constant ties between truck events, Releases and the monitor block."""
    return load_cfg("integer_mine")
