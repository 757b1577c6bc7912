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
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.fixture(scope="session")
def real_map():
    from botlab_b200 import synth
    g = load_golden("real_map")
    return synth.GridSpec(g["cells"], float(g["origin_x"]), float(g["origin_y"]), float(g["meters_per_cell"]),
                          float(g["cells_per_meter"]))


@pytest.fixture(scope="session")
def sensor_golden():
    return load_golden("sensor")


def synth_grid_from_golden(sg):
    from botlab_b200 import synth
    geom = sg["synth_geom"]
    return synth.GridSpec(sg["synth_cells"], float(geom[0]), float(geom[1]), float(geom[2]), float(geom[3]))
