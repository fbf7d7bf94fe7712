import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'multigpu: needs two or more CUDA devices (NCCL); not part of the one-GPU `-m gpu` run')
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.device_count() < 2:                  # multi-GPU tests are deselected (not reported as skips) where they cannot run
        multi = [it for it in items if 'multigpu' in it.keywords]
        if multi:
            config.hook.pytest_deselected(items=multi)
            items[:] = [it for it in items if 'multigpu' not in it.keywords]
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)
