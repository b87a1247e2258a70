import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def synthetic_weights():
    from bugcar_image_segmentation_b200 import weights as W
    with open(os.path.join(ROOT, "pretrained_models", "enet_synthetic_seed42.bcw"), "rb") as f:
        blob = f.read()
    w, nc, eps = W.unpack_flat(blob)
    return blob, w, nc, eps
