import importlib
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def sea():
    """The product package (hyphenated directory name -> importlib)."""
    return importlib.import_module('sea-attention_b200')


def load_golden(name):
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    return {k: g[k] for k in g.files}


def golden_layer(name):
    g = load_golden(name)
    meta = dict(zip(['N', 'H', 'd', 'T', 'k', 'P', 'nbf', 'causal'], g['meta'].tolist()))
    sd = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith('sd.')}
    return g, meta, sd
