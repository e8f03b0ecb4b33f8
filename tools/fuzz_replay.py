"""Randomised differential test on the GPU box: random shapes / densities / hard-site counts / seeds, every
case replayed through the three sweep kernels (one thread per column, per-taxon c/d, large-shape with a random
block size and group-buffer budget) and compared bit for bit with the oracle on the same tape; then three free-running variants against the oracle
reproducing the Philox stream (detmath).

    python tools/fuzz_replay.py [cases=60] [seed=1]      -> gpurun_out/fuzz_replay.json
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
import seriation_b200 as S      # noqa: E402


def oracle_chain(X, hard, seed, burn, samp, manycd):
    o = O.Oracle(X, hard)
    if manycd:
        o.manycd()
    o.source_mt(seed).record(True)
    o.randomize()
    states = []
    for _ in range(burn + samp):
        o.sample()
        states.append(o.state())
    return o.tape(), states


def compare(run, states, burn, samp, manycd):
    bad = []
    run.advance(burn, False).sync()
    for s in range(samp):
        run.advance(1, True).sync()
        got, want = run.state(0), states[burn + s]
        for k in ("a", "b", "pi", "rpi", "t0", "f0", "t1", "f1", "tot"):
            if not np.array_equal(got[k], getattr(want, k)):
                bad.append((s, k))
        if got["slots"] != want.slots or got["loglik"] != want.loglik:
            bad.append((s, "slots/loglik"))
        if manycd:
            c, d = run.cd(0)
            if c.tobytes() != want.c_all.tobytes() or d.tobytes() != want.d_all.tobytes():
                bad.append((s, "cd"))
        elif got["c"] != want.c or got["d"] != want.d:
            bad.append((s, "cd"))
    if run.check() != 0:
        bad.append(("check", run.flags(0)))
    return bad


def run_fuzz(cases=60, seed=1):
    """-> dict(cases, variants_per_case, failures, refused, passed); also used by tests/test_gpu_parity.py"""
    rng = np.random.default_rng(seed)
    failures, log = [], []
    for case in range(cases):
        N = int(rng.integers(2, 300)) if rng.random() < 0.8 else int(rng.integers(300, 700))
        M = int(rng.integers(1, 400))
        dens = float(rng.choice([0.02, 0.1, 0.3, 0.6, 0.95]))
        X = (rng.random((N, M)) < dens).astype(np.uint8)
        if rng.random() < 0.3:
            X[:, rng.integers(0, M)] = 0            # an all-zero column
        if rng.random() < 0.2:
            X[:, rng.integers(0, M)] = 1            # a full column
        nh = int(rng.choice([0, 1, 2, min(N, 7), N - 1, N]))
        hard = np.zeros(N, np.uint8)
        hard[rng.choice(N, size=min(nh, N), replace=False)] = 1
        seed, burn, samp = int(rng.integers(0, 1 << 30)), 2, 3
        ds = S.Dataset.from_bits(X, hard)
        variants = [("small", {}, False), ("manycd", {}, True),
                    ("big", {"SER_FORCE_BIG": str(int(rng.choice([32, 64, 256, 1024]))),
                             "SER_BIG_SMEM_KB": str(int(rng.choice([48, 100, 220]))),
                             "SER_BIG_WARP": str(int(rng.choice([0, 1])))}, False),
                    ("groups", {"SER_SWEEP_GROUPS": str(int(rng.integers(1, 9)))}, bool(rng.integers(0, 2)))]
        tapes = {m: oracle_chain(X, hard, seed, burn, samp, m) for m in (False, True)}
        for name, env, manycd in variants:
            for k, v in env.items():
                os.environ[k] = v
            try:
                tape, states = tapes[manycd]
                run = S.Run(ds, 1, mode=S.MODE_REPLAY, manycd=manycd)
                run.set_tapes([tape]).init()
                bad = compare(run, states, burn, samp, manycd)
                run.close()
            except S.SeriationError as e:            # a shape a variant legitimately refuses
                bad = [] if ("manycd" in str(e) or "shared memory" in str(e)) else [("error", str(e))]
                log.append(dict(case=case, variant=name, refused=str(e)))
            finally:
                for k in env:
                    os.environ.pop(k, None)
            if bad:
                failures.append(dict(case=case, variant=name, N=N, M=M, density=dens, nh=nh, seed=seed, env=env, bad=[str(b) for b in bad[:6]]))
        # free-running mode: the oracle reproduces the GPU's Philox stream bit for bit (detmath)
        for name, env, manycd in (("free", {}, False), ("free-manycd", {}, True), ("free-big", {"SER_FORCE_BIG": "128"}, False)):
            for k, v in env.items():
                os.environ[k] = v
            try:
                fseed, gid = int(rng.integers(0, 1 << 31)), int(rng.integers(0, 70000))
                o = O.Oracle(X, hard)
                if manycd:
                    o.manycd()
                o.source_philox(fseed, gid).detmath(True)
                o.randomize()
                for _ in range(3):
                    o.sample()
                want = o.state()
                run = S.Run(ds, 1, seed=fseed, chain_offset=gid, manycd=manycd)
                run.init().advance(1, False).advance(2, True).sync()
                got = run.state(0)
                bad = [k for k in ("a", "b", "pi", "tot") if not np.array_equal(got[k], getattr(want, k))]
                if got["loglik"] != want.loglik:
                    bad.append("loglik")
                if manycd and run.cd(0)[0].tobytes() != want.c_all.tobytes():
                    bad.append("cd")
                if run.check() != 0:
                    bad.append("check")
                run.close()
            except S.SeriationError as e:
                bad = [] if ("manycd" in str(e) or "shared memory" in str(e)) else [("error", str(e))]
                log.append(dict(case=case, variant=name, refused=str(e)))
            finally:
                for k in env:
                    os.environ.pop(k, None)
            if bad:
                failures.append(dict(case=case, variant=name, N=N, M=M, density=dens, nh=nh, bad=[str(b_) for b_ in bad[:6]]))
    return dict(cases=cases, variants_per_case=7, failures=failures, refused=log, passed=not failures)


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    out = run_fuzz(cases, int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    failures, log = out["failures"], out["refused"]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "fuzz_replay.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(dict(cases=cases, failures=len(failures), refused=len(log), passed=not failures)))
    for fl in failures[:5]:
        print(fl)


if __name__ == "__main__":
    main()
