import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
NOW = ("g10s10", "g10s2", "g5s5", "g2s2")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


from tools.datasets import load_hex_dataset  # noqa: E402,F401


@pytest.fixture(scope="session")
def datasets():
    return {n: load_hex_dataset(n) for n in NOW}


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle as O
    O.build()
    return O


def random_dataset(rng, N, M, nh, density):
    X = (rng.random((N, M)) < density).astype(np.uint8)
    hard = np.zeros(N, np.uint8)
    if nh:
        hard[rng.choice(N, nh, replace=False)] = 1
    return X, hard


EDGE_SHAPES = [  # (N, M, nh, density)
    (2, 3, 0, .5), (3, 2, 1, .5), (5, 4, 0, .4), (8, 5, 7, .3), (8, 5, 8, .3), (8, 5, 6, .3), (31, 7, 3, .2),
    (32, 9, 4, .2), (33, 9, 0, .2), (64, 20, 5, .1), (96, 33, 2, .15), (65, 40, 64, .1), (128, 64, 16, .08),
    (40, 70, 10, .0), (200, 31, 3, .05),
]
