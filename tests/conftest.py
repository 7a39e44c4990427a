import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_checkpoint(S):
    """The shipped ``wind_gnn_{S}.pth`` state_dict (saved from cuda:0 -> map to cpu)."""
    import torch

    return torch.load(os.path.join(GOLDEN, f"wind_gnn_{S}.pth"), map_location="cpu", weights_only=True)


def station_latlon(S):
    """(lat, lon) of the 7- or 34-station set, CSV order (step1:32 drops Enchant 2)."""
    with open(os.path.join(GOLDEN, "coords.json")) as f:
        c = json.load(f)
    idx = list(range(7)) if S == 7 else [i for i, n in enumerate(c["names"]) if n != "Enchant 2 AGCM"]
    return np.array([[c["lat"][i], c["lon"][i]] for i in idx], dtype=np.float64)


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def ckpt7():
    return load_checkpoint(7)


@pytest.fixture(scope="session")
def ckpt34():
    return load_checkpoint(34)
