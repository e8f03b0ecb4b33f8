#!/usr/bin/env python
"""BASELINE.json config 2 at full size: g10s2, 100 chains, 1000 burn-in + 1000 sampling calls of
10 sweeps; tapes and per-sample states recorded from the UNMODIFIED reference (oracle/_ref/ref_mcmc,
MT19937 seeds 0..99, one process per host core), replayed on one B200 and compared bit for bit.

    python tools/config2_full.py [n_chains=100] [burn=1000] [samp=1000] [dataset=g10s2]
Writes gpurun_out/replay_full_<dataset>_<chains>x<sweeps>.json."""
import json
import os
import sys
import tempfile
import time
from multiprocessing import Pool

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tools.datasets import load_hex_dataset, write_txt  # noqa: E402


def ref_chain(args):
    path, seed, burn, samp = args
    with tempfile.TemporaryDirectory() as td:
        dims, states, tape = O.ref_trace(path, burn, samp, td, seed=seed)
    st = states[1 + burn:]
    return dict(tape=tape, a=np.stack([s.a for s in st]), b=np.stack([s.b for s in st]), pi=np.stack([s.pi for s in st]),
                c=np.array([s.c for s in st]), d=np.array([s.d for s in st]), ll=np.array([s.loglik for s in st]),
                tot=states[-1].tot, slots=states[-1].slots)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    burn = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    samp = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
    name = sys.argv[4] if len(sys.argv) > 4 else "g10s2"
    import seriation_b200 as S
    X, hard = load_hex_dataset(name)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, name + ".txt")
        write_txt(path, X, hard)
        t0 = time.perf_counter()
        with Pool(os.cpu_count()) as pool:
            ref = pool.map(ref_chain, [(path, seed, burn, samp) for seed in range(n)], chunksize=1)
        t_ref = time.perf_counter() - t0
    run = S.Run(S.Dataset.from_bits(X, hard), n, mode=S.MODE_REPLAY, store=S.STORE_FULL, max_samples=samp)
    run.set_tapes([r["tape"] for r in ref])
    t0 = time.perf_counter()
    run.init().advance(burn, False).advance(samp, True).sync()
    t_gpu = time.perf_counter() - t0
    bad = run.check()
    mism = dict(a=0, b=0, pi=0, c=0, d=0, loglik_bits=0, totals=0, cursor=0)
    max_ll_rel = 0.0
    for i, r in enumerate(ref):
        g = run.fetch_samples(i)
        mism["a"] += int(np.any(g["a"] != r["a"])); mism["b"] += int(np.any(g["b"] != r["b"])); mism["pi"] += int(np.any(g["pi"] != r["pi"]))
        mism["c"] += int(np.any(g["c"] != r["c"])); mism["d"] += int(np.any(g["d"] != r["d"]))
        mism["loglik_bits"] += int(np.any(g["loglik"] != r["ll"]))
        max_ll_rel = max(max_ll_rel, float(np.max(np.abs(g["loglik"] - r["ll"]) / np.abs(r["ll"]))))
        st = run.state(i)
        mism["totals"] += int(np.any(st["tot"] != r["tot"])); mism["cursor"] += int(st["slots"] != r["slots"] or st["slots"] != r["tape"].size)
    out = dict(dataset=name, chains=n, burn_calls=burn, sample_calls=samp, sweeps_per_chain=10 * (burn + samp),
               tape_slots_total=int(sum(r["tape"].size for r in ref)), reference_seconds=t_ref, host_cores=os.cpu_count(),
               gpu_seconds=t_gpu, inconsistent_chains=bad, chains_with_mismatch=mism, max_loglik_rel_err=max_ll_rel,
               passed=bool(bad == 0 and not any(mism.values())))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "replay_full_%s_%dx%d.json" % (name, n, 10 * (burn + samp))), "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
