import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def grid2562():
    from mpas_regent_b200 import mesh as M
    return M.load_npz(os.path.join(ROOT, "tests", "golden", "x1.2562.grid.npz"), name="x1.2562")


@pytest.fixture(scope="session")
def grid642():
    from mpas_regent_b200 import icosa
    return icosa.make_icosahedral_mesh(642)
