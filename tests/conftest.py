import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a real B200 (run with -m gpu on the GPU box)')
    config.addinivalue_line('markers', 'slow: takes more than a few seconds')


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) when no CUDA device is visible."""
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session', autouse=True)
def _built_library():
    """Make sure libkeynet_b200.so and the C oracle exist (no-ops when they are up to date; nvcc / gcc cross-compile
    without a GPU).  Building the checker is not using it."""
    try:
        from keynet_b200 import build as kb
        kb.build()
        from oracle import keynet_oracle
        keynet_oracle.build()
    except Exception as e:          # report, but let the tests that need the library fail loudly on their own
        print('[conftest] build failed: %s' % e)
    yield
