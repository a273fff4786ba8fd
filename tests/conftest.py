import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden_n32():
    return dict(np.load(GOLDEN / "ref_n32.npz"))


@pytest.fixture(scope="session")
def tables_nu64():
    return dict(np.load(GOLDEN / "ref_tables_nu64.npz"))


@pytest.fixture(scope="session")
def tables_nu150():
    return dict(np.load(GOLDEN / "ref_tables_nu150.npz"))


@pytest.fixture(scope="session")
def oracle():
    from oracle.binding import Oracle
    return Oracle()


def params_of(npz: dict):
    """GhCudaParams from a golden npz (scalars + tables exactly as the reference produced them)."""
    from crime_b200.abi import params_from_dict
    d = {k: (v.item() if np.ndim(v) == 0 else v) for k, v in npz.items()}
    return params_from_dict(d)


def field_err(a, b):
    """max|a-b| / rms(b): the tolerance semantics for fields with zero crossings."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / np.sqrt(np.mean(b * b)))
