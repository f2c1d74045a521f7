import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    np.seterr(all="ignore")


def stats_err_arrays(a, b):
    """ratio-1, cosine-1, relative error: test/test_utils.jl:78-83 of the reference."""
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    na, nb = np.linalg.norm(a), np.linalg.norm(b)
    return na / nb - 1.0, float(a @ b) / (na * nb) - 1.0, np.linalg.norm(a - b) / na


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (d if d > 0 else 1.0)


@pytest.fixture(scope="session")
def has_cuda():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False
