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
    # pytest-timeout registers this itself when it is installed; without the plug-in the mark is inert, not an error
    config.addinivalue_line("markers", "timeout(seconds): multi-rank tests must fail, not hang, when a rank is lost")


def pytest_collection_modifyitems(config, items):
    """A plain ``pytest tests`` on a CPU-only host skips the gpu-marked tests instead of failing in Engine()."""
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (run with -m gpu on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


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
