import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def scene_small():
    """320x240 synthetic scene + keyframe + three frames (shared by CPU and GPU tests)."""
    from tests.helpers import make_case
    return make_case(320, 240, n_frames=3, seed=11)


@pytest.fixture(scope="session")
def scene_vga():
    from tests.helpers import make_case
    return make_case(640, 480, n_frames=2, seed=23)
