"""Per-phase cycle shares of the large-shape sweep kernel on the 1024 x 4096 synthetic matrix.
Needs a timing build:  NVCC_EXTRA=-DSER_PHASE_TIMING sh seriation-in-paleontological-data-using-mcmc_b200/build.sh
(thread 0 of every CTA accumulates clock64() deltas between phase marks; rebuild without the flag afterwards)."""
import sys, ctypes as C; sys.path.insert(0,'.')
import seriation_b200 as S
ds = S.Dataset.synthetic(1024, 4096, 16)
run = S.Run(ds, 296, seed=1, store=S.STORE_PI, max_samples=4)
run.init().advance(1, False).sync()
out = (C.c_ulonglong * 24)()
S.lib().ser_debug_phase_cycles(out)
run.advance(2, True).sync(); ms = run.elapsed_ms(reset=True)
S.lib().ser_debug_phase_cycles(out)
tot = sum(out)
names = ["stage+H", "E postings", "S+L", "D dense / warp batches", "PT pick / wait for the CTA's last batch", "totals", "pi (groups build)", "PT scan", "pi1 (proposal, deltas, decision)", "pi2", "pi3 (proposal, deltas, decision)", "swap", "pi1 accepted: columns, site order, hard tables", "pi3 accepted: columns, site order"]
print("%.1f ms for 296 chains x 20 sweeps = %.0f sweeps/s (instrumented build); cycles summed over the CTAs of a chain, per sweep: %.0f" % (ms, 296 * 20 / (ms * 1e-3), tot / (296 * 20)))
for n, v in zip(names, out): print("%-12s %5.1f%%  %.0f cyc/sweep" % (n, 100.0 * v / tot, v / (296 * 20)))
