"""Loader for the committed NOW-subset fixtures (tests/golden/datasets/*.hex; made by
tools/make_golden.py) shared by tests/, bench.py and __graft_entry__.smoke()."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA_DIR = os.path.join(ROOT, "tests", "golden", "datasets")
NOW = ("g10s10", "g10s2", "g5s5", "g2s2")


def load_hex_dataset(name):
    """-> (X uint8 [N][M], hard uint8 [N])"""
    with open(os.path.join(DATA_DIR, name + ".hex")) as f:
        n, m = (int(t) for t in f.readline().split())
        X = np.zeros((n, m), np.uint8)
        hard = np.zeros(n, np.uint8)
        for i in range(n):
            parts = f.readline().split()
            bits = np.unpackbits(np.frombuffer(bytes.fromhex(parts[0]), np.uint8), bitorder="little")
            X[i] = bits[:m]
            hard[i] = len(parts) > 1 and parts[1] == "*"
    return X, hard


def write_txt(path, X, hard):
    """The reference's Dataset/*.txt layout."""
    with open(path, "w") as f:
        f.write("%d %d\n" % X.shape)
        for i in range(X.shape[0]):
            f.write(" ".join(str(int(v)) for v in X[i]) + (" * " if hard[i] else " ") + "\n")
