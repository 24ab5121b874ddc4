import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def _have_gpu():
    return os.path.exists('/dev/nvidia0') or os.path.exists('/dev/nvidiactl')


def pytest_collection_modifyitems(config, items):
    # GPU tests are skipped only where there is no GPU at all; on a GPU box a missing
    # libb200fem.so makes them FAIL (there is no CPU fallback to hide behind).
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason='no CUDA device in this container')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)
