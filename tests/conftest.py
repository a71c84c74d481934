import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def uniform_noise(n, seed):
    """Zero-mean, unit-variance uniform noise like the reference test input
    ((rand() - 0.5) * sqrt(12), src/psd.rs:604-606), but seeded."""
    import numpy as np
    rng = np.random.default_rng(seed)
    return ((rng.random(n, dtype=np.float32) - np.float32(0.5)) * np.float32(12 ** 0.5)).astype(np.float32)


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding
    binding.lib()
    return binding
