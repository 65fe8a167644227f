import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


@pytest.fixture(scope="session")
def bunny():
    return load_golden("bunny_points")["points"]


@pytest.fixture(scope="session")
def egg_carton():
    return load_golden("egg_carton_points")["points"]


@pytest.fixture(scope="session")
def torus_c1():
    return load_golden("torus_c1_points")["points"]
