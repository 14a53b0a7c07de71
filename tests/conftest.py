"""pytest configuration: registers the ``gpu`` marker and makes the repo root importable."""

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# JAQMC_B200_TC_MIN_WORK is a tuning switch (launch size below which the CUDA-core dense kernel is taken; default 0).
# The parity tests run a handful of walkers and are there to check the tensor-core kernels: pin it.
os.environ.setdefault("JAQMC_B200_TC_MIN_WORK", "0")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
