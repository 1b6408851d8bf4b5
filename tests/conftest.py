import os
import sys

import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.abspath(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def blob():
    from real_robots_b200.blob import read_blob
    return read_blob(os.path.join(ROOT, "real_robots_b200", "assets", "scene.rrb"))
