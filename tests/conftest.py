import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def built():
    """Build the product libraries and the oracle once per session (no-ops when fresh)."""
    from qubatron_b200 import build as qb_build
    paths = qb_build.build_all()
    from oracle import qb_oracle
    qb_oracle.build(ref=True)
    return paths


@pytest.fixture(scope="session")
def scene_c1():
    from qubatron_b200 import scene
    return scene.make_c1()


@pytest.fixture(scope="session")
def scene_random():
    from qubatron_b200 import scene
    return scene.make_random(n_static=30000, n_dynamic=4000, seed=11)
