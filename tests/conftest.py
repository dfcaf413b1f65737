import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def ctx():
    """One device context per test session (GPU tests only)."""
    import tantivy_aggregations_b200 as ta
    c = ta.Context(0)
    yield c
    c.close()
