import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def liver():
    from tests import helpers
    return helpers.load_liver()


@pytest.fixture(scope="session")
def liver_cases(liver):
    """{'cond' | 'nocond': dict(ai, simple, prob, mask)} built with the package's host mirror."""
    from tests import helpers
    return {mode: helpers.liver_case(liver, mode) for mode in ("cond", "nocond")}
