import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
TASKS = ('left', 'straight', 'right')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden_common():
    return dict(np.load(os.path.join(GOLDEN, 'common.npz'), allow_pickle=False))


@pytest.fixture(scope='session')
def golden_task():
    cache = {}

    def _get(task):
        if task not in cache:
            cache[task] = dict(np.load(os.path.join(GOLDEN, 'task_%s.npz' % task), allow_pickle=False))
        return cache[task]
    return _get


def golden_paths(g):
    return [(g['path%d_x' % i], g['path%d_y' % i], g['path%d_phi' % i]) for i in range(3)]
