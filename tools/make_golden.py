#!/usr/bin/env python
"""Generates the committed fixtures under tests/golden/ from the reference tree.

Run in the build container (needs /root/reference and oracle/_ref):
    python tools/make_golden.py

* tests/golden/datasets/<name>.hex   the four NOW subsets (Dataset/*.txt of the reference; data by
  Puolamaki, Fortelius & Mannila, CC BY 2.5), re-encoded as one hex string of packed bits per site
  (+ '*' for hard sites) so the GPU box -- which has no /root/reference -- can run on them.
* tests/golden/ref_*.npz             traces of the UNMODIFIED reference (oracle/_ref/ref_mcmc):
  the recorded draw tape and the full model state after every mcmc_sample() call.
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

REF = "/root/reference/Dataset"
OUT = os.path.join(ROOT, "tests", "golden")


def write_hex(name):
    X, hard = O.load_dataset(f"{REF}/{name}.txt")
    os.makedirs(f"{OUT}/datasets", exist_ok=True)
    with open(f"{OUT}/datasets/{name}.hex", "w") as f:
        f.write("%d %d\n" % X.shape)
        for i in range(X.shape[0]):
            bits = np.packbits(X[i], bitorder="little").tobytes().hex()
            f.write(bits + (" *" if hard[i] else "") + "\n")
    return X, hard


def trace(name, burn, samp, seed, step=False, philox=None, tag="", manycd=False):
    with tempfile.TemporaryDirectory() as td:
        dims, states, tape = O.ref_trace(f"{REF}/{name}.txt", burn, samp, td, seed=seed, step=step, philox=philox, manycd=manycd)
    keys = ("a", "b", "pi", "rpi", "t0", "f0", "t1", "f1", "tot")
    out = {k: np.stack([getattr(s, k) for s in states]).astype(np.int16 if k != "tot" else np.int32) for k in keys}
    out["cdl"] = np.array([[s.c, s.d, s.loglik] for s in states])
    out["kind"] = np.array([s.kind for s in states], np.int8)
    out["ret"] = np.array([s.ret for s in states], np.int32)
    out["slots"] = np.array([s.slots for s in states], np.int64)
    out["tape"] = tape
    out["meta"] = np.array([burn, samp, seed, int(step)], np.int64)
    if manycd:
        out["c_all"] = np.stack([s.c_all for s in states])
        out["d_all"] = np.stack([s.d_all for s in states])
    np.savez_compressed(f"{OUT}/ref_{name}{tag}.npz", **out)
    print(name, tag, "records", len(states), "tape", tape.size)


if __name__ == "__main__":
    O.build()
    for n in ("g10s10", "g10s2", "g5s5", "g2s2"):
        write_hex(n)
    trace("g10s10", 6, 6, seed=42)
    trace("g10s10", 1, 1, seed=7, step=True, tag="_step")
    trace("g10s2", 2, 2, seed=3)
    trace("g5s5", 2, 2, seed=4)
    trace("g2s2", 2, 2, seed=5)
    trace("g10s10", 2, 2, seed=0, philox=(20060206, 17), tag="_philox")
    trace("g10s10", 2, 2, seed=8, tag="_manycd", manycd=True)
