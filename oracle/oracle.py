"""ctypes front end to the ORACLE (test infrastructure only).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module.  It wraps

* ``oracle/liboracle.so``  -- the CPU restatement (``seriation_oracle.c``), and
* ``oracle/_ref/ref_mcmc`` -- the unmodified reference ``mcmc.c`` built against
  the GSL-API shim (present when built in a container that has
  ``/root/reference``; the prebuilt binary travels to the GPU box).

It also restates, in numpy, the analysis side of ``script.py``: ``choose_chains`` :70-99,
``compute_exp_cd`` :102-126, ``compute_exp_ages`` :129-152, ``compute_pair_order_matrix`` :155-189,
``compute_exp_pi`` / ``compute_exp_a`` :230-276 and the three probability maps :306-448.
Parity status of these restatements: PINNED -- ``tests/golden/script_g10s10.npz`` holds the outputs of the
unmodified ``script.py`` (imported here by ``tools/make_golden_script.py``) and
``tests/test_oracle_golden.py`` compares every one of them.
"""
from __future__ import annotations

import ctypes as C
import os
import struct
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REF_BIN = os.path.join(REF_DIR, "ref_mcmc")

_lib = None


def build(force: bool = False) -> None:
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    target = ["oracle", "ref"] if os.path.isdir("/root/reference/C_Implementation") else ["oracle"]
    if force and os.path.exists(LIB_PATH):
        os.remove(LIB_PATH)
    subprocess.run(["make", "-C", HERE, "-s"] + target, check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        vp, i32p, dp, u8p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_uint8)
        L.orc_create.restype = vp
        L.orc_create.argtypes = [C.c_int, C.c_int, u8p, u8p]
        L.orc_free.argtypes = [vp]
        L.orc_source_mt.argtypes = [vp, C.c_ulong]
        L.orc_source_philox.argtypes = [vp, C.c_uint32, C.c_uint32]
        L.orc_source_tape.argtypes = [vp, dp, C.c_size_t]
        L.orc_record.argtypes = [vp, C.c_int]
        L.orc_tape_len.restype = C.c_size_t
        L.orc_tape_len.argtypes = [vp]
        L.orc_tape_copy.argtypes = [vp, dp]
        L.orc_tape_slots.restype = C.c_longlong
        L.orc_tape_slots.argtypes = [vp]
        L.orc_tape_mismatches.restype = C.c_longlong
        L.orc_tape_mismatches.argtypes = [vp]
        L.orc_set_detmath.argtypes = [vp, C.c_int]
        L.orc_set_manycd.argtypes = [vp, C.c_int]
        L.orc_get_cd.argtypes = [vp, dp, dp]
        for name in ("orc_randomize", "orc_recount"):
            getattr(L, name).argtypes = [vp]
            getattr(L, name).restype = None
        for name in ("orc_samplec", "orc_sampled", "orc_sampleab", "orc_samplepi1", "orc_samplepi3",
                     "orc_sweep", "orc_sample", "orc_consistent"):
            getattr(L, name).argtypes = [vp]
            getattr(L, name).restype = C.c_int
        L.orc_samplepi2.argtypes = [vp, C.c_int]
        L.orc_samplepi2.restype = C.c_int
        L.orc_get_state.argtypes = [vp] + [i32p] * 9 + [dp]
        L.orc_set_state.argtypes = [vp, i32p, i32p, i32p, C.c_double, C.c_double]
        L.orc_run.argtypes = [vp, C.c_int, C.c_int, i32p, i32p, i32p, dp, i32p, dp]
        L.orc_margins.argtypes = [vp, dp, dp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
        _lib = L
    return _lib


def _p(arr, ctype):
    return arr.ctypes.data_as(C.POINTER(ctype)) if arr is not None else None


# --------------------------------------------------------------------------- datasets
def parse_dataset_text(text: str):
    """Restates mcmc_readmodel's parsing rules (mcmc.c:339-401): header ``N M``; per row the first
    M '0'/'1' characters (any separators), then a '*' anywhere later marks a hard site."""
    lines = text.split("\n")
    n, m = (int(t) for t in lines[0].split()[:2])
    X = np.zeros((n, m), dtype=np.uint8)
    hard = np.zeros(n, dtype=np.uint8)
    for i in range(n):
        s = lines[1 + i]
        k = j = 0
        while j < m and k < len(s):
            ch = s[k]
            if ch in "01":
                X[i, j] = ch == "1"
                j += 1
            k += 1
        if "*" in s[k:]:
            hard[i] = 1
    return X, hard


def load_dataset(path: str):
    with open(path) as f:
        return parse_dataset_text(f.read())


def format_dataset(X: np.ndarray, hard: np.ndarray) -> str:
    """Write a matrix in the reference's .txt layout (Dataset/*.txt)."""
    out = ["%d %d" % X.shape]
    for i in range(X.shape[0]):
        out.append(" ".join(str(int(v)) for v in X[i]) + (" * " if hard[i] else " "))
    return "\n".join(out) + "\n"


# --------------------------------------------------------------------------- state snapshots
@dataclass
class State:
    kind: int
    ret: int
    slots: int
    a: np.ndarray
    b: np.ndarray
    pi: np.ndarray
    rpi: np.ndarray
    t0: np.ndarray
    f0: np.ndarray
    t1: np.ndarray
    f1: np.ndarray
    tot: np.ndarray
    c: float
    d: float
    loglik: float
    c_all: np.ndarray = None   # per-taxon c, d (manycd)
    d_all: np.ndarray = None

    def same_ints(self, o: "State") -> bool:
        return all(np.array_equal(getattr(self, k), getattr(o, k))
                   for k in ("a", "b", "pi", "rpi", "t0", "f0", "t1", "f1", "tot"))

    def same_bits(self, o: "State") -> bool:
        ok = self.same_ints(o) and struct.pack("3d", self.c, self.d, self.loglik) == \
            struct.pack("3d", o.c, o.d, o.loglik)
        if ok and self.c_all is not None and o.c_all is not None:
            ok = self.c_all.tobytes() == o.c_all.tobytes() and self.d_all.tobytes() == o.d_all.tobytes()
        return ok


def read_dump(path: str):
    """Parse a dump written by ``ref_mcmc trace`` (format in oracle/ref_harness.c)."""
    buf = open(path, "rb").read()
    magic, N, M, nh = struct.unpack_from("<4i", buf, 0)
    assert magic in (0x5345524D, 0x5345524E), "bad dump magic"
    manycd = magic == 0x5345524E
    off = 16
    rec_ints = 2 * M + 2 * N + 4 * M + 4
    out = []
    while off < len(buf):
        kind, ret, slots = struct.unpack_from("<iiq", buf, off)
        off += 16
        ints = np.frombuffer(buf, dtype="<i4", count=rec_ints, offset=off).copy()
        off += 4 * rec_ints
        c, d, ll = struct.unpack_from("<3d", buf, off)
        off += 24
        c_all = d_all = None
        if manycd:
            c_all = np.frombuffer(buf, dtype="<f8", count=M, offset=off).copy()
            d_all = np.frombuffer(buf, dtype="<f8", count=M, offset=off + 8 * M).copy()
            off += 16 * M
        o = 0
        fields = []
        for ln in (M, M, N, N, M, M, M, M, 4):
            fields.append(ints[o:o + ln])
            o += ln
        out.append(State(kind, ret, slots, *fields, c, d, ll, c_all, d_all))
    return (N, M, nh), out


# --------------------------------------------------------------------------- the restatement
class Oracle:
    """One chain of the CPU restatement."""

    def __init__(self, X: np.ndarray, hard: np.ndarray):
        self.X = np.ascontiguousarray(X, dtype=np.uint8)
        self.hard = np.ascontiguousarray(hard, dtype=np.uint8)
        self.N, self.M = self.X.shape
        self._h = lib().orc_create(self.N, self.M, _p(self.X, C.c_uint8), _p(self.hard, C.c_uint8))
        self._tape = None

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_free(self._h)
            self._h = None

    # draw sources
    def source_mt(self, seed: int):
        lib().orc_source_mt(self._h, seed)
        return self

    def source_philox(self, seed: int, chain: int):
        lib().orc_source_philox(self._h, seed, chain)
        return self

    def source_tape(self, tape: np.ndarray):
        self._tape = np.ascontiguousarray(tape, dtype=np.float64)
        lib().orc_source_tape(self._h, _p(self._tape, C.c_double), self._tape.size)
        return self

    def record(self, on: bool = True):
        lib().orc_record(self._h, int(on))
        return self

    def tape(self) -> np.ndarray:
        n = lib().orc_tape_len(self._h)
        out = np.empty(n, dtype=np.float64)
        if n:
            lib().orc_tape_copy(self._h, _p(out, C.c_double))
        return out

    def detmath(self, on: bool = True):
        lib().orc_set_detmath(self._h, int(on))
        return self

    def manycd(self, on: bool = True):
        """per-taxon c, d; call BEFORE choosing the draw source"""
        lib().orc_set_manycd(self._h, int(on))
        self._manycd = bool(on)
        return self

    def cd(self):
        c, d = np.empty(self.M), np.empty(self.M)
        lib().orc_get_cd(self._h, _p(c, C.c_double), _p(d, C.c_double))
        return c, d

    @property
    def slots(self) -> int:
        return lib().orc_tape_slots(self._h)

    @property
    def tape_mismatches(self) -> int:
        return lib().orc_tape_mismatches(self._h)

    # sampler
    def randomize(self):
        lib().orc_randomize(self._h)
        return self

    def samplec(self): return lib().orc_samplec(self._h)
    def sampled(self): return lib().orc_sampled(self._h)
    def sampleab(self): return lib().orc_sampleab(self._h)
    def samplepi1(self): return lib().orc_samplepi1(self._h)
    def samplepi2(self, swap: int): return lib().orc_samplepi2(self._h, swap)
    def samplepi3(self): return lib().orc_samplepi3(self._h)
    def sweep(self): return lib().orc_sweep(self._h)
    def sample(self): return lib().orc_sample(self._h)
    def consistent(self) -> int: return lib().orc_consistent(self._h)

    def state(self, kind: int = -1, ret: int = 0) -> State:
        N, M = self.N, self.M
        a, b, t0, f0, t1, f1 = (np.empty(M, np.int32) for _ in range(6))
        pi, rpi = np.empty(N, np.int32), np.empty(N, np.int32)
        tot, cdl = np.empty(4, np.int32), np.empty(3, np.float64)
        lib().orc_get_state(self._h, *(_p(v, C.c_int32) for v in (a, b, pi, rpi, t0, f0, t1, f1, tot)),
                            _p(cdl, C.c_double))
        c_all, d_all = self.cd() if getattr(self, "_manycd", False) else (None, None)
        return State(kind, ret, self.slots, a, b, pi, rpi, t0, f0, t1, f1, tot, *cdl, c_all, d_all)

    def set_state(self, a, b, pi, c: float, d: float):
        a, b, pi = (np.ascontiguousarray(v, dtype=np.int32) for v in (a, b, pi))
        lib().orc_set_state(self._h, _p(a, C.c_int32), _p(b, C.c_int32), _p(pi, C.c_int32), c, d)
        return self

    def run(self, burn_calls: int, sample_calls: int):
        """burn + sampling mcmc_sample() calls; returns per-sample arrays and the exp_data sums."""
        N, M, S = self.N, self.M, sample_calls
        a, b = np.empty((S, M), np.int32), np.empty((S, M), np.int32)
        pi = np.empty((S, N), np.int32)
        cdl, counts, sums = np.empty((S, 3)), np.empty((S, 4), np.int32), np.empty(3)
        lib().orc_run(self._h, burn_calls, S, _p(a, C.c_int32), _p(b, C.c_int32), _p(pi, C.c_int32),
                      _p(cdl, C.c_double), _p(counts, C.c_int32), _p(sums, C.c_double))
        return dict(a=a, b=b, pi=pi, c=cdl[:, 0].copy(), d=cdl[:, 1].copy(), loglik=cdl[:, 2].copy(),
                    counts=counts, sums=sums)

    def margins(self):
        mp, ma = C.c_double(), C.c_double()
        nd, npr = C.c_longlong(), C.c_longlong()
        lib().orc_margins(self._h, C.byref(mp), C.byref(ma), C.byref(nd), C.byref(npr))
        return dict(min_pick=mp.value, min_accept=ma.value, n_degenerate=nd.value, n_proposals=npr.value)


# --------------------------------------------------------------------------- the real reference
def ref_available() -> bool:
    return os.access(REF_BIN, os.X_OK)


def ref_trace(dataset_path: str, burn_calls: int, sample_calls: int, workdir: str, *, seed: int = 0,
              philox=None, step: bool = False, tape_in: str | None = None, binary: str = REF_BIN, manycd: bool = False):
    """Run the unmodified reference under the harness; returns (dims, states, tape)."""
    dump = os.path.join(workdir, "ref.dump")
    tape = os.path.join(workdir, "ref.tape")
    env = dict(os.environ, SER_TAPE_OUT=tape)
    env.pop("SER_TAPE_IN", None)
    env.pop("SER_RNG", None)
    if tape_in:
        env["SER_TAPE_IN"] = tape_in
    elif philox is not None:
        env.update(SER_RNG="philox", SER_SEED=str(philox[0]), SER_CHAIN=str(philox[1]))
    else:
        env["GSL_RNG_SEED"] = str(seed)
    cmd = [binary, "trace", dataset_path, str(burn_calls), str(sample_calls), dump] + (["step"] if step else []) + \
        (["manycd"] if manycd else [])
    subprocess.run(cmd, check=True, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    dims, states = read_dump(dump)
    return dims, states, np.fromfile(tape, dtype="<f8")


# --------------------------------------------------------------------------- cross-chain steps
def choose_chains(exp_neg_loglik: np.ndarray, k: int):
    """script.py:70-99: chains with min-sigma < x < min+sigma (population sigma over ALL chains,
    strict), the k smallest of them, returned as sorted chain ids."""
    x = np.asarray(exp_neg_loglik, dtype=np.float64)
    lo = x.min()
    sd = np.std(x)
    cand = [i for i in range(len(x)) if lo - sd < x[i] < lo + sd]
    cand.sort(key=lambda i: (x[i], i))
    return sorted(cand[:k])


def pair_order_counts(pi_samples: np.ndarray) -> np.ndarray:
    """script.py:178-189 summed over samples: cnt[i][j] = #{t: pi_t(i) < pi_t(j)}, diagonal = -T."""
    pi = np.asarray(pi_samples)
    T, N = pi.shape
    cnt = np.zeros((N, N), dtype=np.int64)
    for t in range(T):
        cnt += pi[t][:, None] < pi[t][None, :]
    cnt[np.arange(N), np.arange(N)] = -T
    return cnt


def pair_order_matrix(per_chain_counts, chains_selected: int, faithful: bool = True) -> np.ndarray:
    """script.py:155-175.  With ``faithful`` the reference's carry-over is kept: po_matrix_chain is
    not reset between chains, so chain k starts from chain k-1's already /1000-scaled matrix."""
    N = per_chain_counts[0].shape[0]
    po = np.zeros((N, N))
    carry = np.zeros((N, N))
    for cnt in per_chain_counts:
        acc = (carry if faithful else 0.0) + cnt.astype(np.float64)
        acc = acc / 1000
        po += acc
        carry = acc
    return po / chains_selected


def exp_ages(pi_samples_per_chain, chains_selected: int) -> float:
    """script.py:129-152 verbatim in numpy: per chain sum_t pearsonr(pi_t, arange(N)) / 1000, summed, / chains_selected"""
    total = 0.0
    for pis in pi_samples_per_chain:
        pis = np.asarray(pis, dtype=np.float64)
        ref = np.arange(pis.shape[1], dtype=np.float64)
        s = 0.0
        for p in pis:
            s += np.corrcoef(p, ref)[0, 1]
        total += s / 1000
    return total / chains_selected


def exp_pi(pi_samples_per_chain, chains_selected: int) -> np.ndarray:
    """script.py:230-252 verbatim: pi_sum reset inside the loop (:243), pi_sum_chain never reset"""
    pi_sum = 0
    pi_sum_chain = np.zeros(np.asarray(pi_samples_per_chain[0]).shape[1])
    for pis in pi_samples_per_chain:
        pi_sum = 0
        for p in np.asarray(pis):
            pi_sum_chain += p
        pi_sum_chain /= 1000
        pi_sum += pi_sum_chain
    return pi_sum / chains_selected


def exp_a(a_samples_per_chain, chains_selected: int) -> np.ndarray:
    """script.py:255-276 verbatim: a_sum accumulates, a_sum_chain never reset"""
    a_sum = np.zeros(np.asarray(a_samples_per_chain[0]).shape[1])
    a_sum_chain = np.zeros_like(a_sum)
    for a_s in a_samples_per_chain:
        for a in np.asarray(a_s):
            a_sum_chain += a
        a_sum_chain /= 1000
        a_sum += a_sum_chain
    return a_sum / chains_selected


def exp_cd(c_samples_per_chain, d_samples_per_chain, chains_selected: int):
    """script.py:102-126: the file carries exp(c) printed with %.14f; per chain sum / 1000, summed, / chains_selected"""
    c_tot = d_tot = 0.0
    for cs, ds_ in zip(c_samples_per_chain, d_samples_per_chain):
        c_sum = d_sum = 0
        for c, d in zip(cs, ds_):
            c_sum += float("%.14f" % np.exp(c))
            d_sum += float("%.14f" % np.exp(d))
        c_tot += c_sum / 1000
        d_tot += d_sum / 1000
    return c_tot / chains_selected, d_tot / chains_selected


def _shuffle_like_script(X_sum, exp_pi_v, exp_a_v):
    """script.py:339-353: rows by the rank of each site in exp_pi, columns by argsort(exp_a)"""
    rpi = np.argsort(exp_pi_v)
    idx = np.empty_like(rpi)
    idx[rpi] = np.arange(len(rpi))
    X_sum = X_sum[idx, :]
    ra = np.argsort(exp_a_v)
    out = np.zeros_like(X_sum)
    for i, taxon in enumerate(ra):
        out[:, i] = X_sum[:, taxon]
    return out


def _interval_maps(a_samples_per_chain, b_samples_per_chain, N, chains_selected, cell):
    """common loop of script.py:306-448: X_sum_chain is divided by 1000 but never reset between chains"""
    M = np.asarray(a_samples_per_chain[0]).shape[1]
    j = np.arange(N)[:, None]
    X_sum_chain = np.zeros((N, M))
    X_sum = np.zeros((N, M))
    for a_s, b_s in zip(a_samples_per_chain, b_samples_per_chain):
        for a, b in zip(np.asarray(a_s), np.asarray(b_s)):
            alive = (j >= a[None, :]) & (j <= b[None, :])   # closed at b (script.py:329), unlike the sampler's [a, b)
            X_sum_chain += cell(alive)
        X_sum_chain /= 1000
        X_sum += X_sum_chain
    return X_sum / chains_selected


def alive_matrix(a_spc, b_spc, pi_spc, chains_selected: int) -> np.ndarray:
    """plot_taxa_occurence_probability_matrix, script.py:306-353 (the returned matrix; no plot)"""
    N = np.asarray(pi_spc[0]).shape[1]
    X = _interval_maps(a_spc, b_spc, N, chains_selected, lambda alive: alive.astype(np.float64))
    return _shuffle_like_script(X, exp_pi(pi_spc, chains_selected), exp_a(a_spc, chains_selected))


def false_taxa_matrix(a_spc, b_spc, pi_spc, chains_selected: int) -> np.ndarray:
    """plot_false_taxa_occurence_probability, script.py:356-403"""
    N = np.asarray(pi_spc[0]).shape[1]
    X = _interval_maps(a_spc, b_spc, N, chains_selected, lambda alive: (~alive).astype(np.float64))
    return _shuffle_like_script(X, exp_pi(pi_spc, chains_selected), exp_a(a_spc, chains_selected))


def false_ones_matrix(a_spc, b_spc, pi_spc, chains_selected: int, Xdata: np.ndarray) -> np.ndarray:
    """plot_false_ones_probability, script.py:406-448: X[j][i] is read in FILE order while a, b are
    positions -- the reference mixes the two index spaces; kept"""
    N = np.asarray(pi_spc[0]).shape[1]
    ones = np.asarray(Xdata) == 1
    X = _interval_maps(a_spc, b_spc, N, chains_selected, lambda alive: (ones & ~alive).astype(np.float64))
    return _shuffle_like_script(X, exp_pi(pi_spc, chains_selected), exp_a(a_spc, chains_selected))
