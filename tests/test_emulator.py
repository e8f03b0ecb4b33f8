"""CPU emulation of the kernel's choreography (tests/emul/chain_emul.cpp) against the oracle.

The emulator compiles the very building blocks the CUDA kernel uses (csrc/ser_chain_core.h) for
the host, so the bit-level logic -- range popcounts, rank/select over the hard mask, the pi3 mask
construction, the Gibbs walk, the degenerate-delta rule -- is verified without a GPU.  Replay
mode: integers bit-exact, c/d bit-exact, loglik within 1e-9 relative.  Every Gibbs step is also run the way the large-shape
kernel's warp batches run it (a column's 1 .. 32 lanes own contiguous item chunks: chunk-fused weights, chunk-relative sums,
scan of the chunk totals, search + pick in the first chunk that reaches the target): the fused weights equal the dense ones bit
for bit and the picked boundary equals the sequential inverse CDF's."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

from conftest import EDGE_SHAPES, NOW, ROOT, load_hex_dataset, random_dataset

EMUL_DIR = os.path.join(ROOT, "tests", "emul")


@pytest.fixture(scope="module")
def emul():
    so = os.path.join(EMUL_DIR, "libchain_emul.so")
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", so,
                    os.path.join(EMUL_DIR, "chain_emul.cpp")], check=True)
    L = C.CDLL(so)
    u8p, i32p, dp = C.POINTER(C.c_uint8), C.POINTER(C.c_int32), C.POINTER(C.c_double)
    L.emul_create.restype = C.c_void_p
    L.emul_create.argtypes = [C.c_int, C.c_int, u8p, u8p] + [C.c_double] * 5
    L.emul_free.argtypes = [C.c_void_p]
    L.emul_set_tape.argtypes = [C.c_void_p, dp, C.c_size_t]
    L.emul_randomize.argtypes = [C.c_void_p]
    L.emul_sweeps.argtypes = [C.c_void_p, C.c_int]
    L.emul_chunk_stats.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
    L.emul_get_state.argtypes = [C.c_void_p] + [i32p] * 9 + [dp, C.POINTER(C.c_longlong)]
    return L


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _emul_state(O, L, h, N, M):
    a, b, t0, f0, t1, f1 = (np.empty(M, np.int32) for _ in range(6))
    pi, rpi = np.empty(N, np.int32), np.empty(N, np.int32)
    tot, cdl, sl = np.empty(4, np.int32), np.empty(3), C.c_longlong()
    L.emul_get_state(h, *(_p(v, C.c_int32) for v in (a, b, pi, rpi, t0, f0, t1, f1, tot)), _p(cdl, C.c_double), C.byref(sl))
    return O.State(-1, 0, sl.value, a, b, pi, rpi, t0, f0, t1, f1, tot, *cdl)


def _run_case(O, L, X, hard, seed, sweeps):
    N, M = X.shape
    o = O.Oracle(X, hard).source_mt(seed).record(True)
    o.randomize()
    want = [o.state()]
    for _ in range(sweeps):
        o.sweep()
        want.append(o.state())
    tape = o.tape()
    c0, d0 = math.log(.01), math.log(.3)
    e = L.emul_create(N, M, _p(X, C.c_uint8), _p(hard, C.c_uint8), c0, math.log(1. - math.exp(c0)), d0,
                      math.log(1. - math.exp(d0)), math.exp(-32.236191301916641))
    try:
        L.emul_set_tape(e, _p(tape, C.c_double), tape.size)
        L.emul_randomize(e)
        for k in range(sweeps + 1):
            if k:
                L.emul_sweeps(e, 1)
            s, r = _emul_state(O, L, e, N, M), want[k]
            assert s.same_ints(r), (k, [f for f in ("a", "b", "pi", "rpi", "t0", "f0", "t1", "f1", "tot")
                                        if not np.array_equal(getattr(s, f), getattr(r, f))])
            assert s.slots == r.slots and s.c == r.c and s.d == r.d, k
            assert abs(s.loglik - r.loglik) <= 1e-9 * abs(r.loglik), k
        st3 = (C.c_longlong * 3)()
        L.emul_chunk_stats(e, st3)
        # the warp batches' chunked passes (1 .. 32 lanes per column) beside the sequential inverse CDF: same boundary, one finder lane
        assert st3[0] > 0 and st3[1] == 0 and st3[2] == 0, list(st3)
    finally:
        L.emul_free(e)
    assert o.consistent() == 0


@pytest.mark.parametrize("shape", EDGE_SHAPES)
def test_emulator_edge_shapes(oracle_mod, emul, shape):
    rng = np.random.default_rng(hash(shape) & 0xffff)
    X, hard = random_dataset(rng, *shape)
    for seed in (1, 2):
        _run_case(oracle_mod, emul, X, hard, seed, 200)


@pytest.mark.parametrize("name,sweeps", [("g10s10", 1500), ("g5s5", 200), ("g10s2", 150), ("g2s2", 150)])
def test_emulator_now_subsets(oracle_mod, emul, name, sweeps):
    X, hard = load_hex_dataset(name)
    _run_case(oracle_mod, emul, X, hard, 9, sweeps)


def test_emulator_random_shapes_and_densities(oracle_mod, emul):
    """The fuzz generator of tools/fuzz_replay.py on the CPU: random shapes, densities up to 95 %, all-zero /
    full columns, 0..N hard sites -- the kernel's building blocks against the oracle, sweep by sweep."""
    rng = np.random.default_rng(2024)
    for _ in range(25):
        N, M = int(rng.integers(2, 160)), int(rng.integers(1, 120))
        dens = float(rng.choice([0.02, 0.1, 0.3, 0.6, 0.95]))
        X = (rng.random((N, M)) < dens).astype(np.uint8)
        if rng.random() < 0.3:
            X[:, rng.integers(0, M)] = 0
        if rng.random() < 0.2:
            X[:, rng.integers(0, M)] = 1
        nh = int(rng.choice([0, 1, 2, min(N, 7), N - 1, N]))
        hard = np.zeros(N, np.uint8)
        hard[rng.choice(N, size=min(nh, N), replace=False)] = 1
        _run_case(oracle_mod, emul, X, hard, int(rng.integers(0, 1 << 30)), 12)


def test_coarse_prefix_tables_equal_per_word_tables(emul):
    """the large-shape kernel keeps one prefix count per 2^G words of a column (ser_pre_at<G>, G = 2 in the product): random moves
    (site move, reversal, adjacent swap, window permutation) and range queries on a coarse and a per-word table must agree"""
    emul.emul_coarse_pre_selftest.argtypes = [C.c_uint64, C.c_int, C.c_int]
    emul.emul_coarse_pre_selftest.restype = C.c_int
    for seed, N in enumerate((2, 5, 31, 32, 33, 64, 95, 127, 128, 129, 200, 526, 1000, 1024, 2047, 2048)):
        assert emul.emul_coarse_pre_selftest(seed + 1, N, 300) == 0, N
