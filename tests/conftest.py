import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope='session')
def oracle_models():
    from oracle.physics import OracleModel, load_model_json
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = OracleModel(load_model_json('%s_%s' % (name, name)))
        return cache[name]
    return get
