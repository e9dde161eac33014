import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

RBF1 = [0.34608543, 1.0, 0.34608543]
RBF2 = [0.08263808, 0.53616077, 1.0, 0.53616077, 0.08263808]
MAT15_2 = [0.15233751, 0.50067621, 1.0, 0.50067621, 0.15233751]
MAT15_3 = [0.08435782, 0.24239115, 0.60311586, 1.0, 0.60311586, 0.24239115, 0.08435782]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def make_inputs(N, d, L, seed=0, dist="randn", scale=1.0):
    """Seeded synthetic inputs: x ~ N(0, I) (or U[0,1]) fp32, V ~ N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, d, generator=g) if dist == "randn" else torch.rand(N, d, generator=g)
    v = torch.randn(N, L, generator=g)
    return (x * scale).contiguous(), v.contiguous()


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.int32) if a.dtype == np.float32 else a


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.lib()
    return o


@pytest.fixture(scope="session")
def sg():
    import simplex_gp_b200
    return simplex_gp_b200
