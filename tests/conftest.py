"""pytest configuration: markers and shared fixtures.

``-m "not gpu"``  oracle vs golden vectors, host logic, C-ABI symbol/export checks, gloo N>1 logic.
``-m gpu``        parity tests proper: CUDA path (through the C-ABI) vs oracle / goldens on a B200.
"""
import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_files(prefix):
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))


def load_golden(path):
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def has_cuda():
    import torch
    return torch.cuda.is_available()
