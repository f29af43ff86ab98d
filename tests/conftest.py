import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def models():
    """The reference's 210 fixture models: {scale: {'GRAPHS': ..., 'THETAS': {str(j): [10 lists]}}}"""
    return json.load(open(os.path.join(GOLDEN, 'models.json')))


@pytest.fixture(scope='session')
def aer_counts():
    """The reference's stored Aer histograms (res_*/result_simulation.json)."""
    return json.load(open(os.path.join(GOLDEN, 'aer_counts.json')))


def all_models(models):
    for scale in ('0.1', '0.25', '0.5'):
        for j, C in enumerate(models[scale]['GRAPHS']):
            for i, th in enumerate(models[scale]['THETAS'][str(j)]):
                yield scale, j, i, C, th


@pytest.fixture(scope='session')
def native_built():
    from qcmrf_b200 import build
    return build.build_native()


def has_cuda():
    try:
        from qcmrf_b200 import _native
        return _native.device_count() > 0
    except Exception:
        return False
