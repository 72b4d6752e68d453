import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import harness
    gpu = None
    for item in items:
        if "gpu" in item.keywords:
            if gpu is None:
                gpu = harness.have_gpu()
            if not gpu:
                item.add_marker(pytest.mark.skip(reason="no CUDA device"))
