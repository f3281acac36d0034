import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(ROOT, "tests", "golden")
SHIPPED_WEIGHTS = os.path.join(ROOT, "livingscenes_b200", "_weights", "shipped_fp32.pt")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def state_dict_for(tag):
    """'random' -> seeded weights (always available); 'shipped' -> the extracted checkpoint or skip."""
    from oracle import restatement as R

    if tag == "random":
        return R.random_state_dict(0)
    if os.path.exists(SHIPPED_WEIGHTS):
        return torch.load(SHIPPED_WEIGHTS, map_location="cpu", weights_only=True)
    from oracle import ref_loader

    if ref_loader.available() and ref_loader.checkpoint_available():
        return ref_loader.shipped_state_dict()
    pytest.skip("shipped checkpoint not available (run __graft_entry__.build() where /root/reference exists)")


def relerr(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.fixture(scope="session")
def oracle_R():
    from oracle import restatement as R

    return R
