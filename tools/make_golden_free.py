"""Reference result of a FULL free-running analysis, for north_star's statistical check (3):

    python tools/make_golden_free.py        # ~10 min on 8 cores; writes tests/golden/ref_free_g10s10.npz

Runs the UNMODIFIED reference end to end on g10s10 exactly as script.py does: 100 chains of
`mcmc <i> < g10s10.txt` (1000 burn-in + 1000 sampling calls of 10 sweeps, GSL_RNG_SEED = i, MT19937 behind
the GSL-API shim), then the unmodified script.py's choose_chains(8), compute_pair_order_matrix,
compute_exp_cd and compute_exp_ages over the Chains/ directory.  The GPU's free-running mode uses a
different random stream, so the comparison in tests/test_gpu_parity.py is statistical: E[-logL] of the
selected chains within one reference sigma, pair-order matrix close to the reference's.
Runs in this container only (needs /root/reference and oracle/_ref/ref_mcmc).
"""
import importlib.util
import os
import subprocess
import sys
import tempfile
import types
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
N_CHAINS, K = 100, 8


def main():
    from conftest import load_hex_dataset
    from tools.datasets import write_txt
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    spec = importlib.util.spec_from_file_location("ref_script", "/root/reference/script.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "ref_mcmc")
    X, hard = load_hex_dataset("g10s10")
    N, M = X.shape
    with tempfile.TemporaryDirectory() as td:
        ds = os.path.join(td, "g10s10.txt")
        write_txt(ds, X, hard)
        for i in range(N_CHAINS):
            os.makedirs(os.path.join(td, "Chains", "chain_%02d" % i))

        def one(i):
            with open(ds) as f:
                subprocess.run([ref_bin, "cli", str(i)], stdin=f, cwd=td, env=dict(os.environ, GSL_RNG_SEED=str(i)),
                               check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        with ThreadPoolExecutor(os.cpu_count() or 1) as ex:
            list(ex.map(one, range(N_CHAINS)))
        cwd = os.getcwd()
        os.chdir(td)
        try:
            e = np.array([float(open("Chains/chain_%02d/exp_data.csv" % i).read().split("\n")[1].split(",")[0]) for i in range(N_CHAINS)])
            chosen = ref.choose_chains(K)
            out = dict(e_negloglik=e, chosen=np.array(chosen), exp_cd=np.array(ref.compute_exp_cd(chosen, K)),
                       exp_ages=np.array(ref.compute_exp_ages(chosen, K, N)),
                       po=ref.compute_pair_order_matrix(chosen, K, N).astype(np.float32))
        finally:
            os.chdir(cwd)
    path = os.path.join(ROOT, "tests", "golden", "ref_free_g10s10.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "chosen", chosen, "E[-logL] min %.2f sigma %.2f" % (e.min(), e.std()), "exp_cd", out["exp_cd"], "corr", out["exp_ages"])


if __name__ == "__main__":
    main()
