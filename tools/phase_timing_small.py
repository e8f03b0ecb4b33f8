"""Per-phase cycle shares of ser_sweep_kernel (one thread per column) from its own cycle counters.
Needs a timing build next to the product library (thread 0 of every CTA accumulates clock64() deltas between marks):
    SER_OUT=libseriation_b200_phase.so NVCC_EXTRA=-DSER_PHASE_TIMING sh seriation-in-paleontological-data-using-mcmc_b200/build.sh
    SERIATION_B200_LIB=$PWD/seriation-in-paleontological-data-using-mcmc_b200/libseriation_b200_phase.so python tools/phase_timing_small.py [dataset] [chains]
The proposal marks charge the time since the previous mark to the previous proposal's kind."""
import ctypes as C
import sys
sys.path.insert(0, '.')
import seriation_b200 as S
from tools.datasets import load_hex_dataset
name = sys.argv[1] if len(sys.argv) > 1 else "g2s2"
chains = int(sys.argv[2]) if len(sys.argv) > 2 else 4440
ds = S.Dataset.from_bits(*load_hex_dataset(name))
run = S.Run(ds, chains, seed=1, store=S.STORE_PI, max_samples=20)
run.init().advance(100, False).sync()
out = (C.c_ulonglong * 24)()
S.lib().ser_debug_phase_cycles(out)
run.elapsed_ms(reset=True)
run.advance(20, True).sync()
ms = run.elapsed_ms(reset=True)
S.lib().ser_debug_phase_cycles(out)
n = chains * 200
tot = sum(out)
print("%s: %d chains x 200 sweeps in %.1f ms = %.0f sweeps/s (instrumented build); %.0f cycles per sweep and CTA" % (name, chains, ms, n / (ms * 1e-3), tot / n))
names = {8: "draws, c/d, H table", 9: "postings (expand_ones)", 10: "step geometry (owner)", 19: "maximum + cached log-weights", 11: "dense item weights", 12: "scan + item of the uniform", 18: "pick inside the run (owner)",
         13: "totals / loglik", 14: "pi1 proposals", 15: "pi2 proposals", 16: "pi3 proposals (+ sweep tail)", 17: "swap proposal"}
print("  items evaluated per sweep %.0f, above the LOGEPSILON floor %.1f %%" % (out[20] / n, 100.0 * out[21] / max(1, out[20])))
tot -= out[20] + out[21]
for i, nm in names.items():
    print("  %-30s %5.1f %%  %8.0f cycles/sweep" % (nm, 100.0 * out[i] / tot, out[i] / n))
