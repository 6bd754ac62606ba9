import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def problem():
    from trifocal_pose_estimation_using_improved_gpuhc_b200 import fixtures
    return fixtures.load_problem()


@pytest.fixture(scope="session")
def ransac0():
    from trifocal_pose_estimation_using_improved_gpuhc_b200 import fixtures
    return fixtures.load_ransac(0)


@pytest.fixture(scope="session")
def oracle(problem):
    from oracle.pyoracle import Oracle
    return Oracle(problem)
