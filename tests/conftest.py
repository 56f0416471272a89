import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def engine_factory():
    """Engine constructor; skips cleanly on CPU-only hosts (the tests that use it are marked gpu)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from lagrangian_microbes_b200.engine import Engine
    made = []

    def make(**kw):
        e = Engine(**kw)
        made.append(e)
        return e
    yield make
    for e in made:
        try:
            e.close()
        except Exception:      # a faulted context must not turn the whole session into an error
            pass
