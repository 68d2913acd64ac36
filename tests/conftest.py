import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def lj():
    """The product package (directory name has hyphens -> importlib)."""
    return importlib.import_module("lennard-jones-live-simulator_b200")


@pytest.fixture(scope="session")
def ljlib(lj):
    return lj.load_library()
