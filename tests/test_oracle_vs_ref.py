"""Live comparison of the restatement with the unmodified reference binary (oracle/_ref/ref_mcmc).
Runs wherever the prebuilt binary is present (it travels to the GPU box); otherwise skipped --
tests/test_oracle_golden.py holds the committed equivalent."""
import os
import tempfile

import numpy as np
import pytest

from conftest import NOW, load_hex_dataset


@pytest.fixture(scope="module")
def O(oracle_mod):
    if not oracle_mod.ref_available():
        pytest.skip("oracle/_ref/ref_mcmc not built (needs /root/reference at build time)")
    return oracle_mod


def _dataset_file(O, name, td):
    X, hard = load_hex_dataset(name)
    path = os.path.join(td, name + ".txt")
    with open(path, "w") as f:
        f.write(O.format_dataset(X, hard))
    return X, hard, path


@pytest.mark.parametrize("name,calls", [("g10s10", 12), ("g10s2", 2), ("g5s5", 3), ("g2s2", 2)])
def test_restatement_bit_exact_vs_live_reference(O, name, calls):
    with tempfile.TemporaryDirectory() as td:
        X, hard, path = _dataset_file(O, name, td)
        dims, states, tape = O.ref_trace(path, calls // 2, calls - calls // 2, td, seed=1234 + calls)
    o = O.Oracle(X, hard).source_tape(tape)
    o.randomize()
    assert o.state().same_bits(states[0])
    for r in range(1, len(states)):
        o.sample()
        s = o.state()
        assert s.same_bits(states[r]) and s.slots == states[r].slots, r
    assert o.tape_mismatches == 0


def test_restatement_mt_stream_equals_shim_mt_stream(O):
    """Same MT19937 seed -> the restatement's own generator reproduces the shim's tape."""
    with tempfile.TemporaryDirectory() as td:
        X, hard, path = _dataset_file(O, "g10s10", td)
        dims, states, tape = O.ref_trace(path, 2, 2, td, seed=99)
    o = O.Oracle(X, hard).source_mt(99).record(True)
    o.randomize()
    for r in range(1, len(states)):
        o.sample()
        assert o.state().same_bits(states[r])
    assert np.array_equal(o.tape(), tape)


def test_reference_step_trace(O):
    with tempfile.TemporaryDirectory() as td:
        X, hard, path = _dataset_file(O, "g5s5", td)
        dims, states, tape = O.ref_trace(path, 1, 1, td, seed=5, step=True)
    o = O.Oracle(X, hard).source_tape(tape)
    o.randomize()
    step = {10: o.samplec, 11: o.sampled, 12: o.sampleab, 13: lambda: o.samplepi2(1), 14: o.samplepi1,
            15: lambda: o.samplepi2(0), 16: o.samplepi3}
    for r in range(1, len(states)):
        if states[r].kind == 1:
            continue
        assert step[states[r].kind]() == states[r].ret
        assert o.state().same_bits(states[r]), r
