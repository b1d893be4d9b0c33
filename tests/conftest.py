import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # float(loss) on a tensor that still carries a graph is what the assertions mean to do
    config.addinivalue_line("filterwarnings", "ignore:Converting a tensor with requires_grad=True to a scalar:UserWarning")


@pytest.fixture(scope="session")
def golden():
    import json

    import numpy as np

    d = os.path.join(ROOT, "tests", "golden")
    arrays = np.load(os.path.join(d, "reference_golden.npz"))
    meta = json.load(open(os.path.join(d, "reference_golden.json")))
    return arrays, meta
