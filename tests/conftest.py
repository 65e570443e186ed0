import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def synth_np(w, h, seed=1234, frame=0, **kw):
    from kinectdepthmapenhancement_b200 import synth
    d, c = synth.rgbd_frame(w, h, seed, frame, **kw)
    return d.numpy(), c.numpy()


def rule_active_mask(depth, mean64, window, sigma_d, margin_mm=1.0):
    """See oracle.guard_active_mask (the definition bench.py's parity block uses too)."""
    import oracle
    return oracle.guard_active_mask(depth, mean64, window, sigma_d, margin_mm)
