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
