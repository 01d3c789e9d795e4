import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def golden_input(g, **gen):
    """Regenerate a fixture's input from its seed and check it against the stored digest."""
    import hashlib
    from tetraear_b200 import synth
    x = synth.carrier_iq(int(g["n_samples"]), **gen)
    assert hashlib.sha256(np.ascontiguousarray(x).view(np.uint8)).hexdigest() == str(g["input_sha256"]), \
        "synthetic generator drifted from the golden fixtures"
    return x


@pytest.fixture(scope="session")
def gpu_processor():
    from tetraear_b200.processor import SignalProcessor
    sp = SignalProcessor(2.4e6)
    yield sp
    sp.close()
