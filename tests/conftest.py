import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"))
    return load


def synth_weights(I, F, P, seed):
    """Random weights with the statistics of the trained LENS models."""
    rng = np.random.default_rng(seed)
    kind = rng.random((F, I))
    Wf = np.where(kind < 0.30, rng.exponential(0.147, (F, I)),
                  np.where(kind < 0.83, -rng.exponential(0.089, (F, I)), 0.0))
    Wf = np.clip(Wf, -4.7, 1.05).astype(np.float32)
    Wo = np.clip(rng.normal(0.0, 0.0143, (P, F)), -0.11, 0.055)
    Wo = np.where(np.abs(Wo) < 1e-6, 1e-6, Wo).astype(np.float32)
    return Wf, Wo


def synth_pooled(B, Q, I, seed):
    """Pooled pixel counts: geometric with mean ~8, ~43 % zeros, wrapped to u8."""
    rng = np.random.default_rng(seed)
    v = rng.geometric(1.0 / 15.0, (B, Q, I)) - 1
    v = np.where(rng.random((B, Q, I)) < 0.40, 0, v)
    return (v % 256).astype(np.uint8)
