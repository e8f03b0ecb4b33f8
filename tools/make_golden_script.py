"""Golden vectors for the cross-chain / posterior-summary side, produced by the UNMODIFIED reference
script.py (imported from /root/reference with matplotlib/seaborn stubbed -- they only plot).

    python tools/make_golden_script.py        # writes tests/golden/script_g10s10.npz

Input chains: g10s10, the structured Philox stream (seed SEED, chains 0..N_CHAINS-1) run through the
oracle in detmath mode -- the GPU's free-running mode reproduces exactly these chains, so the -m gpu
tests can compare the product with what script.py computes, without shipping tapes.  The Chains/
directory is written in the reference's own format (mcmc.c:60-92), then script.py's functions are
called with the working directory there.  Runs in this container only (needs /root/reference).
"""
import importlib.util
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

SEED, N_CHAINS, BURN, SAMP, K = 4242, 6, 5, 40, 3


def write_chain_dir(d, res, M):
    """chain_data.csv / exp_data.csv the way mcmc_save_chain and print_exp_data lay them out"""
    os.makedirs(d)
    with open(os.path.join(d, "chain_data.csv"), "w") as f:
        for s in range(len(res["c"])):
            ec, ed = "%.14f " % np.exp(res["c"][s]), "%.14f " % np.exp(res["d"][s])
            f.write("".join("%d " % v for v in res["a"][s]) + "," + "".join("%d " % v for v in res["b"][s]) + "," +
                    "".join("%d " % v for v in res["pi"][s]) + "," + ec * M + "," + ed * M + ",%.14f\n" % res["loglik"][s])
    with open(os.path.join(d, "exp_data.csv"), "w") as f:
        f.write("exp_loglik,exp_c,exp_d\n%.14f,%.14f,%.14f" % tuple(res["sums"] / 1000))


def main():
    import oracle as O
    from conftest import load_hex_dataset
    from tools.datasets import write_txt
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    spec = importlib.util.spec_from_file_location("ref_script", "/root/reference/script.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)

    X, hard = load_hex_dataset("g10s10")
    N, M = X.shape
    out = dict(meta=np.array([SEED, N_CHAINS, BURN, SAMP, K]))
    with tempfile.TemporaryDirectory() as td:
        for i in range(N_CHAINS):
            o = O.Oracle(X, hard).source_philox(SEED, i).detmath(True)
            o.randomize()
            res = o.run(BURN, SAMP)
            write_chain_dir(os.path.join(td, "Chains", "chain_%02d" % i), res, M)
            out["e_negloglik_%d" % i] = np.array(res["sums"][0] / SAMP)
        write_txt(os.path.join(td, "g10s10.txt"), X, hard)
        cwd = os.getcwd()
        os.chdir(td)
        try:
            chosen = ref.choose_chains(K)
            out["chosen"] = np.array(chosen)
            out["exp_cd"] = np.array(ref.compute_exp_cd(chosen, K))
            out["exp_ages"] = np.array(ref.compute_exp_ages(chosen, K, N))
            out["po"] = ref.compute_pair_order_matrix(chosen, K, N)
            out["exp_pi"] = np.array(ref.compute_exp_pi(chosen, N, K))
            out["exp_a"] = np.array(ref.compute_exp_a(chosen, K, M))
            out["alive"] = ref.plot_taxa_occurence_probability_matrix(chosen, K, N, M)
            out["false_taxa"] = ref.plot_false_taxa_occurence_probability(chosen, K, N, M)
            out["false_ones"] = ref.plot_false_ones_probability(chosen, K, "g10s10.txt", N, M)
        finally:
            os.chdir(cwd)
    path = os.path.join(ROOT, "tests", "golden", "script_g10s10.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: np.asarray(v).shape for k, v in out.items() if not k.startswith("e_")})


if __name__ == "__main__":
    main()
