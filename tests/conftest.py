import os
import sys
import json
import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, 'gast-net-3dposeestimation_b200')
for p in (PKG, REPO):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(REPO, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    """`-m gpu` tests need a CUDA device and the built library; without them they are skipped (not failed),
    so that a plain `pytest tests` is green on a CPU-only machine."""
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    lib = os.path.join(PKG, 'csrc', 'libgast_b200.so')
    if have and os.path.exists(lib):
        return
    why = 'no CUDA device' if not have else 'libgast_b200.so is not built'
    skip = pytest.mark.skip(reason='gpu test: ' + why)
    for it in items:
        if 'gpu' in it.keywords:
            it.add_marker(skip)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    d = {k: z[k] for k in z.files if k != 'meta'}
    if 'meta' in z.files:
        d['meta'] = json.loads(str(z['meta']))
    return d


def golden_names(prefix):
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith('.npz') and f.startswith(prefix))


@pytest.fixture(scope='session')
def golden():
    return load_golden
