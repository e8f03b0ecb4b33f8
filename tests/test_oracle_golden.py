"""The oracle restatement against traces of the UNMODIFIED reference (committed fixtures made by
tools/make_golden.py from oracle/_ref/ref_mcmc).  Bit-exact on every integer and on c, d, loglik."""
import os
import struct

import numpy as np
import pytest

from conftest import GOLDEN, NOW, load_hex_dataset


def _bits(x):
    return struct.pack("d", float(x))


def _cmp(state, g, r):
    for k in ("a", "b", "pi", "rpi", "t0", "f0", "t1", "f1", "tot"):
        assert np.array_equal(getattr(state, k), g[k][r].astype(np.int32)), (k, r)
    assert state.slots == int(g["slots"][r]), r
    for got, want in zip((state.c, state.d, state.loglik), g["cdl"][r]):
        assert _bits(got) == _bits(want), r


@pytest.mark.parametrize("name", NOW)
def test_restatement_matches_reference_trace(oracle_mod, name):
    g = np.load(os.path.join(GOLDEN, f"ref_{name}.npz"))
    X, hard = load_hex_dataset(name)
    o = oracle_mod.Oracle(X, hard).source_tape(g["tape"])
    o.randomize()
    _cmp(o.state(), g, 0)
    for r in range(1, len(g["kind"])):
        o.sample()  # (the reference's return value accumulates in statics, mcmc.c:220 -- not compared)
        _cmp(o.state(), g, r)
    assert o.slots == g["tape"].size
    assert o.tape_mismatches == 0 and o.consistent() == 0


def test_restatement_matches_reference_per_subsampler(oracle_mod):
    g = np.load(os.path.join(GOLDEN, "ref_g10s10_step.npz"))
    X, hard = load_hex_dataset("g10s10")
    o = oracle_mod.Oracle(X, hard).source_tape(g["tape"])
    o.randomize()
    _cmp(o.state(), g, 0)
    step = {10: o.samplec, 11: o.sampled, 12: o.sampleab, 13: lambda: o.samplepi2(1), 14: o.samplepi1,
            15: lambda: o.samplepi2(0), 16: o.samplepi3}
    for r in range(1, len(g["kind"])):
        k = int(g["kind"][r])
        if k == 1:
            continue
        assert step[k]() == int(g["ret"][r]), r
        _cmp(o.state(), g, r)


def test_restatement_matches_reference_on_philox_stream(oracle_mod):
    """The reference driven by the structured Philox stream through the shim == the restatement
    generating the same stream itself (no tape involved)."""
    g = np.load(os.path.join(GOLDEN, "ref_g10s10_philox.npz"))
    X, hard = load_hex_dataset("g10s10")
    o = oracle_mod.Oracle(X, hard).source_philox(20060206, 17).record(True)
    o.randomize()
    _cmp(o.state(), g, 0)
    for r in range(1, len(g["kind"])):
        o.sample()
        _cmp(o.state(), g, r)
    assert np.array_equal(o.tape(), g["tape"])


def test_restatement_matches_reference_manycd(oracle_mod):
    """per-taxon c, d (manycd = 1, mcmc.c:777-785, :807-815): trace of the unmodified reference"""
    g = np.load(os.path.join(GOLDEN, "ref_g10s10_manycd.npz"))
    X, hard = load_hex_dataset("g10s10")
    o = oracle_mod.Oracle(X, hard).manycd().source_tape(g["tape"])
    o.randomize()
    _cmp(o.state(), g, 0)
    for r in range(1, len(g["kind"])):
        o.sample()
        s = o.state()
        _cmp(s, g, r)
        assert s.c_all.tobytes() == g["c_all"][r].tobytes() and s.d_all.tobytes() == g["d_all"][r].tobytes()
    assert o.slots == g["tape"].size and len(set(g["c_all"][-1])) > 100


def test_philox_known_answers():
    """Philox4x32-10 known-answer vectors (Random123 kat_vectors)."""
    import ctypes as C
    import subprocess
    import tempfile
    src = r'''
    #include <stdio.h>
    #include "%s/seriation-in-paleontological-data-using-mcmc_b200/csrc/ser_detmath.h"
    int main(void){ uint32_t o[4];
      ser_philox4x32_10(0,0,0,0,0,0,o); printf("%%08x %%08x %%08x %%08x\n",o[0],o[1],o[2],o[3]);
      ser_philox4x32_10(0xffffffffu,0xffffffffu,0xffffffffu,0xffffffffu,0xffffffffu,0xffffffffu,o); printf("%%08x %%08x %%08x %%08x\n",o[0],o[1],o[2],o[3]);
      ser_philox4x32_10(0x243f6a88u,0x85a308d3u,0x13198a2eu,0x03707344u,0xa4093822u,0x299f31d0u,o); printf("%%08x %%08x %%08x %%08x\n",o[0],o[1],o[2],o[3]);
      return 0; }''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with tempfile.TemporaryDirectory() as td:
        with open(td + "/k.c", "w") as f:
            f.write(src)
        subprocess.run(["gcc", "-std=gnu99", "-O1", "-o", td + "/k", td + "/k.c", "-lm"], check=True)
        out = subprocess.run([td + "/k"], check=True, capture_output=True, text=True).stdout.split("\n")
    assert out[0] == "6627e8d5 e169c58d bc57ac4c 9b00dbd8"
    assert out[1] == "408f276d 41c83b0e a20bc7c6 6d5451fd"
    assert out[2] == "d16cfe09 94fdcceb 5001e420 24126ea1"


def test_script_restatements_match_unmodified_script_py(oracle_mod):
    """The numpy restatements of script.py's cross-chain / posterior steps against outputs of the
    UNMODIFIED script.py (tools/make_golden_script.py imported it here and ran its functions over a
    Chains/ directory holding exactly these chains)."""
    g = np.load(os.path.join(GOLDEN, "script_g10s10.npz"))
    seed, n_chains, burn, samp, k = (int(v) for v in g["meta"])
    X, hard = load_hex_dataset("g10s10")
    res = []
    for i in range(n_chains):
        o = oracle_mod.Oracle(X, hard).source_philox(seed, i).detmath(True)
        o.randomize()
        res.append(o.run(burn, samp))
    e = np.array([float("%.14f" % (r["sums"][0] / 1000)) for r in res])   # what exp_data.csv carries
    chosen = oracle_mod.choose_chains(e, k)
    assert chosen == g["chosen"].tolist()
    pis, a_s, b_s = ([res[c][key] for c in chosen] for key in ("pi", "a", "b"))
    ec, ed = oracle_mod.exp_cd([res[c]["c"] for c in chosen], [res[c]["d"] for c in chosen], k)
    assert abs(ec - g["exp_cd"][0]) < 1e-15 and abs(ed - g["exp_cd"][1]) < 1e-15
    assert abs(oracle_mod.exp_ages(pis, k) - float(g["exp_ages"])) < 1e-13
    po = oracle_mod.pair_order_matrix([oracle_mod.pair_order_counts(p) for p in pis], k)
    assert np.allclose(po, g["po"], rtol=0, atol=1e-15)
    assert np.allclose(oracle_mod.exp_pi(pis, k), g["exp_pi"], rtol=0, atol=1e-13)
    assert np.allclose(oracle_mod.exp_a(a_s, k), g["exp_a"], rtol=0, atol=1e-13)
    assert np.allclose(oracle_mod.alive_matrix(a_s, b_s, pis, k), g["alive"], rtol=0, atol=1e-15)
    assert np.allclose(oracle_mod.false_taxa_matrix(a_s, b_s, pis, k), g["false_taxa"], rtol=0, atol=1e-15)
    assert np.allclose(oracle_mod.false_ones_matrix(a_s, b_s, pis, k, X), g["false_ones"], rtol=0, atol=1e-15)
    assert g["alive"].max() > 0 and g["false_ones"].max() > 0


def test_reference_pipeline_golden_reproduces_report_table1(oracle_mod):
    """tests/golden/ref_free_g10s10.npz: 100 full-length chains of the unmodified mcmc.c + the unmodified
    script.py analysis (tools/make_golden_free.py).  The published Table 1 row for g10s10 is E[c] 0.0119,
    E[d] 0.5127, corr 0.94 -- the reference built here (GSL-API shim, MT19937) lands on it, and the
    restated choose_chains picks the same chains."""
    g = np.load(os.path.join(GOLDEN, "ref_free_g10s10.npz"))
    assert oracle_mod.choose_chains(g["e_negloglik"], 8) == g["chosen"].tolist()
    assert abs(g["exp_cd"][0] - 0.0119) < 0.0005 and abs(g["exp_cd"][1] - 0.5127) < 0.005 and abs(float(g["exp_ages"]) - 0.94) < 0.005
    po = g["po"].astype(np.float64)
    off = ~np.eye(124, dtype=bool)
    assert np.allclose((po + po.T)[off], 1.0, atol=2e-3)   # 8 chains x 1000 samples / 1000 / 8 (+ the carry-over quirk's 1e-3)


@pytest.mark.parametrize("name", ["ref_g10s10", "ref_g10s2", "ref_g5s5", "ref_g2s2", "ref_g10s10_philox", "ref_g10s10_manycd"])
def test_golden_tapes_keep_every_float_decision_far_from_its_boundary(oracle_mod, name):
    """SURVEY section 7 "decisions must match; floats need not": the GPU evaluates the Gibbs weights in another order
    (closed-form runs) and with its own exp, so its picks equal the reference's only while no draw lands within
    rounding distance of a CDF step or of an accept threshold.  Audit the committed tapes: the smallest distance
    of a uniform from the neighbouring CDF steps (mcmc_randompick, mcmc.c:910-913; probabilities sum to 1) and
    the smallest |delta - log u| of the MH tails (mcmc.c:1261/:1441/:1636) stay many orders above 1e-12."""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    X, hard = load_hex_dataset("g10s10" if "g10s10" in name else name[4:])
    o = oracle_mod.Oracle(X, hard)
    if name.endswith("manycd"):
        o.manycd()
    o.source_tape(g["tape"])
    o.randomize()
    for _ in range(1, len(g["kind"])):
        o.sample()
    assert o.slots == g["tape"].size and o.tape_mismatches == 0
    m = o.margins()
    assert m["n_proposals"] >= 16 * 10 * (len(g["kind"]) - 1) - 5 * 10 * (len(g["kind"]) - 1)   # hard-site refusals skip the tail
    assert m["min_pick"] > 1e-9, m
    assert m["min_accept"] > 1e-9, m


def test_restatement_matches_reference_on_the_config5_matrix(oracle_mod):
    """tests/golden/ref_synthetic_1024x4096.npz (tools/make_golden_big.py): the UNMODIFIED reference (MAXS raised so
    it can read 8 192-character rows) on the 1024 x 4096 synthetic matrix of BASELINE.json config 5.  CPU suite:
    chain 0 through its first two calls (20 sweeps); the GPU suite replays all 8 chains x 100 sweeps."""
    import zlib
    import seriation_b200 as S
    g = np.load(os.path.join(GOLDEN, "ref_synthetic_1024x4096.npz"))
    X, hard = S.Dataset.synthetic(1024, 4096, 16).arrays()
    assert [zlib.crc32(X.tobytes()), zlib.crc32(hard.tobytes())] == g["x_crc"].tolist()   # the generator is frozen
    o = oracle_mod.Oracle(X, hard).source_mt(int(g["seeds"][0])).record(True)
    o.randomize()
    for r in range(3):
        if r:
            o.sample()
        st = o.state()
        for k in ("a", "b", "pi"):
            assert np.array_equal(getattr(st, k), g[k][0][r].astype(np.int32)), (k, r)
        assert np.array_equal(st.tot, g["tot"][0][r]) and st.slots == int(g["slots"][0][r])
        assert (st.c, st.d, st.loglik) == tuple(g["cdl"][0][r]), r
