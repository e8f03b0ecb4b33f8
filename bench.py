#!/usr/bin/env python
"""bench.py -- aggregate MCMC sweeps/s across chains (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
    python bench.py --impl reference --steps K --warmup W    (the reference's CPU path)

One *step* = one whole job over the batch of chains on this rank: randomised start (init kernel),
burn-in sweeps, sampling sweeps with thinned-sample streaming, E[-logL] per chain, the
within-one-sigma chain selection and the pair-order counts of the selected chains.  Multi-GPU:
chains are sharded by global chain id (no data-path traffic); the only collectives are one
all-gather of E[-logL] and one all-reduce of the k x N x N pair-order counts (NCCL via
torch.distributed), exactly the end-of-run exchange the path has.

Prints ONE JSON line (rank 0).  `value` times the job with all inputs resident in HBM;
`e2e` times the same job through the host-buffer C ABI (dataset upload, run creation, results
copied back to host).  `roofline` reports the sweep kernel against the SM-local FP64 ceiling
measured live by ser_microbench (the path is neither HBM- nor tensor-bound);
`cpu_baseline` times the unmodified reference (oracle/_ref) on the host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic work per sweep (SURVEY.md section 8d closed forms): canonical cells C, smem bytes B = C/8,
# fp64 candidate weights F = M(N+2) -> 10 F flops, HBM bytes per thinned sample
ALGO = {
    "g10s10": dict(N=124, M=139, C=74967, B=9371, F=17514, hbm_sample=828, k=8),
    "g10s2": dict(N=501, M=139, C=301352, B=37669, F=69917, hbm_sample=1582, k=2),
    "g5s5": dict(N=273, M=202, C=238360, B=29795, F=55550, hbm_sample=1378, k=2),
    "g2s2": dict(N=526, M=296, C=673301, B=84163, F=156288, hbm_sample=2260, k=4),
    "synthetic": dict(N=1024, M=4096, C=18149376, B=2268672, F=4202496, hbm_sample=18456, k=2),
}


def load_dataset(name):
    """(X, hard): a NOW subset from the committed fixtures, or the deterministic 1024 x 4096 synthetic
    matrix of BASELINE.json config 5 (generator: ser_dataset_synthetic)"""
    if name == "synthetic":
        import seriation_b200 as S
        return S.Dataset.synthetic(1024, 4096, 16).arrays()
    from tools.datasets import load_hex_dataset
    return load_hex_dataset(name)
METRIC = "mcmc_sweeps_per_s_aggregate"
UNIT = "sweeps/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dataset", default="g2s2", choices=sorted(ALGO))
    ap.add_argument("--chains", type=int, default=0, help="chains per GPU (0: 16384; synthetic: 8192)")
    ap.add_argument("--burn-calls", type=int, default=0, help="0: 5 (synthetic: 1)")
    ap.add_argument("--sample-calls", type=int, default=0, help="0: 5 (synthetic: 1)")
    ap.add_argument("--no-synthetic-probe", action="store_true")
    ap.add_argument("--cpu-calls", type=int, default=0, help="mcmc_sample() calls per CPU process (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------ CPU reference arm
def cpu_reference_run(dataset, calls, procs):
    """`procs` concurrent processes of the reference sampler (script.py's Pool, :60-62), each doing
    `calls` mcmc_sample() calls (= 10 sweeps each) on `dataset`.  Returns (sweeps/s aggregate, kind)."""
    from oracle import oracle as O
    from tools.datasets import write_txt
    X, hard = load_dataset(dataset)
    ref_bin = O.REF_BIN + ("_big" if X.shape[1] > 900 else "")  # rows > 1999 chars need the MAXS-patched build
    if O.ref_available() and os.access(ref_bin, os.X_OK):
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, dataset + ".txt")
            write_txt(path, X, hard)
            t0 = time.perf_counter()
            ps = [subprocess.Popen([ref_bin, "bench", path, str(calls)], env=dict(os.environ, GSL_RNG_SEED=str(i + 1)),
                                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for i in range(procs)]
            rcs = [p.wait() for p in ps]
            dt = time.perf_counter() - t0
        if any(rcs):
            raise RuntimeError("reference binary failed: %r" % rcs)
        return procs * calls * 10 / dt, "reference"
    # no prebuilt reference binary: time the restatement (one process per core)
    code = ("import sys; sys.path.insert(0, %r); from oracle import oracle as O; import bench;"
            "X,h=bench.load_dataset(%r); o=O.Oracle(X,h).source_mt(int(sys.argv[1])); o.randomize();"
            "[o.sample() for _ in range(%d)]" % (ROOT, dataset, calls))
    O.build()
    t0 = time.perf_counter()
    ps = [subprocess.Popen([sys.executable, "-c", code, str(i + 1)]) for i in range(procs)]
    rcs = [p.wait() for p in ps]
    dt = time.perf_counter() - t0
    if any(rcs):
        raise RuntimeError("oracle port failed: %r" % rcs)
    return procs * calls * 10 / dt, "port"


def auto_cpu_calls(dataset):
    # ~10 s of CPU work per process: the survey measured 24-26 ns per canonical cell
    per_sweep = ALGO[dataset]["C"] * 25e-9
    return max(1, int(10.0 / (per_sweep * 10)))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    procs = os.cpu_count() or 1
    calls = args.cpu_calls or auto_cpu_calls(args.dataset)
    for _ in range(args.warmup):
        cpu_reference_run(args.dataset, max(1, calls // 8), procs)
    t0 = time.perf_counter()
    vals = [cpu_reference_run(args.dataset, calls, procs) for _ in range(args.steps)]
    dt = time.perf_counter() - t0
    value = sum(v for v, _ in vals) / len(vals)
    kind = vals[0][1]
    sample = "%d processes x %d mcmc_sample() calls (%d sweeps each) of %s per step" % (procs, calls, calls * 10, args.dataset)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "NOW subset %s (committed fixture of the reference's Dataset/)" % args.dataset,
        "config": {"workload": "%s, reference C sampler (mcmc.c, -O2, GSL-API shim) on host cores" % args.dataset,
                   "sweeps_per_step": procs * calls * 10},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------ the B200 arm
def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import seriation_b200 as S

    big = args.dataset == "synthetic"
    args.chains = args.chains or (8192 if big else 16384)
    args.burn_calls = args.burn_calls or (1 if big else 5)
    args.sample_calls = args.sample_calls or (1 if big else 5)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    S.lib()  # fails loudly if the CUDA library is missing

    algo = ALGO[args.dataset]
    X, hard = load_dataset(args.dataset)
    N, M, k = X.shape[0], X.shape[1], algo["k"]
    n_local, n_total = args.chains, args.chains * world
    calls = args.burn_calls + args.sample_calls
    sweeps_per_step = n_total * calls * 10
    ds = S.Dataset.from_bits(X, hard)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    d_e_local = torch.empty(n_local, dtype=torch.float64, device=dev)
    d_e_all = torch.empty(n_total, dtype=torch.float64, device=dev)
    d_chosen = torch.empty(k, dtype=torch.int32, device=dev)
    d_info = torch.empty(3, dtype=torch.float64, device=dev)
    d_counts = torch.zeros((k, N, N), dtype=torch.int32, device=dev)
    stream_ptr = torch.cuda.current_stream().cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    launches = 0
    sweep_ms = []

    def cross_chain(run):
        """E[-logL] -> (all-gather) -> selection -> PO counts -> (all-reduce); returns #launches."""
        n = 0
        run.chain_stats_device(d_e_local.data_ptr()); n += 1
        run.sync()
        if world > 1:
            dist.all_gather_into_tensor(d_e_all, d_e_local)
            src = d_e_all
        else:
            src = d_e_local
        S.select_chains_device(src.data_ptr(), n_total, k, d_chosen.data_ptr(), d_info.data_ptr(), local, stream_ptr); n += 1
        d_counts.zero_()
        torch.cuda.current_stream().synchronize()
        run.po_counts_device(d_chosen.data_ptr(), k, d_counts.data_ptr()); n += 1
        run.sync()
        if world > 1:
            dist.all_reduce(d_counts)
        return n

    def step_resident(run):
        """inputs resident: the run object (dataset bits, state buffers) already lives in HBM"""
        n0 = run.kernel_launches()
        run.elapsed_ms(reset=True)
        run.init().advance(args.burn_calls, False).advance(args.sample_calls, True)
        sweep_ms.append(run.elapsed_ms(reset=True))
        cross_chain(run)
        return run.kernel_launches() - n0 + 1  # + the selection kernel (launched outside the run object)

    def step_e2e():
        """host buffers in, host results out, through the C ABI"""
        ds_h = S.Dataset.from_bits(X, hard)                       # host -> library (bit-packed on upload)
        run = S.Run(ds_h, n_local, mode=S.MODE_FREE, seed=20060206, chain_offset=rank * n_local, store=S.STORE_PI,
                    max_samples=args.sample_calls, device=local)
        run.init().advance(args.burn_calls, False).advance(args.sample_calls, True)
        st = run.chain_stats()                                    # D2H: per-chain scalars
        # host buffers -> (all-gather) -> choose_chains -> PO counts (D2H k x N x N) -> (all-reduce) -> PO matrix
        chosen, po = S.cross_chain_distributed(st["e_negloglik"], k, run.po_counts, N)
        run.close()
        return float(po[0, 1])

    run = S.Run(ds, n_local, mode=S.MODE_FREE, seed=20060206, chain_offset=rank * n_local, store=S.STORE_PI,
                max_samples=args.sample_calls, device=local)
    for _ in range(args.warmup):
        step_resident(run)
        flush.zero_()
    sweep_ms.clear()
    clocks = ClockSampler(local)
    barrier()
    if rank == 0:
        clocks.start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        launches += step_resident(run)
        flush.zero_()                                             # L2 flush between timed iterations
    barrier()
    dt = time.perf_counter() - t0
    clk = clocks.stop() if rank == 0 else None
    run.close()

    # end-to-end arm
    for _ in range(args.warmup):
        step_e2e()
    barrier()
    t1 = time.perf_counter()
    e2e_steps = []
    for _ in range(args.steps):
        ts = time.perf_counter()
        step_e2e()
        e2e_steps.append(time.perf_counter() - ts)
    barrier()
    dt_e2e = time.perf_counter() - t1
    if rank == 0:
        print("e2e step times (ms): " + " ".join("%.1f" % (1e3 * v) for v in e2e_steps), file=sys.stderr)

    t = torch.tensor([dt, dt_e2e, sum(sweep_ms) / max(1, len(sweep_ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt, dt_e2e, sweep_kernel_ms = (float(v) for v in t.cpu())

    if rank == 0:
        value = sweeps_per_step * args.steps / dt
        e2e_value = sweeps_per_step * args.steps / dt_e2e
        # roofline of the dominant kernel (ser_sweep_kernel + init, CUDA events on the run's stream)
        mb = S.microbench(local)
        kernel_sweeps_per_s = n_local * calls * 10 / (sweep_kernel_ms * 1e-3)
        achieved = 10.0 * algo["F"] * kernel_sweeps_per_s / 1e12
        # DRAM bytes per sweep of the sweep kernel from the committed `ncu --set full` captures
        # (dram__bytes_read.sum + dram__bytes_write.sum over the chain-sweeps of the profiled launch),
        # scaled to one of this run's two sweep launches per step
        ncu_dram_per_sweep = {"g2s2": (10.899968e6 + 414.976e3) / 40960, "synthetic": (2.160332e9 + 1.837035e9) / 2960}
        traffic = traffic_src = None
        if args.dataset in ncu_dram_per_sweep:
            traffic = ncu_dram_per_sweep[args.dataset] * n_local * calls * 10 / 2
            traffic_src = ("profiles/r01/sweep_%sv7_ncu_raw_selected.txt: %.0f B of DRAM traffic per chain-sweep x the chain-sweeps "
                           "of one sweep launch" % ("big_" if big else "", ncu_dram_per_sweep[args.dataset]))
        roofline = {
            "bound": "fp64", "achieved": achieved, "peak": mb["fp64_tflops"], "unit": "TFLOP/s",
            "frac": achieved / mb["fp64_tflops"], "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": "ser_microbench fp64 FMA on this GPU, measured live (MEASURED_PEAKS.json holds only HBM and bf16)",
            "kernel": "ser_sweep_kernel_big" if big else "ser_sweep_kernel", "kernel_ms_per_step": sweep_kernel_ms,
            "kernel_sweeps_per_s_per_gpu": kernel_sweeps_per_s,
            "smem": {"achieved_gbs": algo["B"] * kernel_sweeps_per_s / 1e9, "peak_gbs": mb["lds_gbs"],
                     "frac": algo["B"] * kernel_sweeps_per_s / 1e9 / mb["lds_gbs"]},
            "int_popc": {"achieved_gops": algo["C"] / 32 * kernel_sweeps_per_s / 1e9, "peak_gops": mb["popc_gops"],
                         "frac": algo["C"] / 32 * kernel_sweeps_per_s / 1e9 / mb["popc_gops"]},
            "hbm": {"achieved_gbs": algo["hbm_sample"] / 10 * kernel_sweeps_per_s / 1e9, "peak_gbs": _hbm_peak(),
                    "frac": algo["hbm_sample"] / 10 * kernel_sweeps_per_s / 1e9 / _hbm_peak()},
        }
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            procs = os.cpu_count() or 1
            ccalls = args.cpu_calls or auto_cpu_calls(args.dataset)
            v, kind = cpu_reference_run(args.dataset, ccalls, procs)
            cpu = {"value": v, "unit": UNIT, "cores": procs, "kind": kind,
                   "sample": "%d processes x %d sweeps of %s (unmodified mcmc.c, -O2, GSL-API shim)" % (procs, ccalls * 10, args.dataset)}
        also = None
        if world == 1 and not big and not args.no_synthetic_probe:
            # BASELINE.json's metric also names the 1024 x 4096 synthetic matrix: a short probe of its
            # sweep kernel (large-shape path), CUDA-event timed, so both shapes appear in one line
            sds = S.Dataset.synthetic(1024, 4096, 16)
            srun = S.Run(sds, 1184, seed=20060206, store=S.STORE_PI, max_samples=1, device=local)
            srun.init().advance(1, False).sync()
            srun.elapsed_ms(reset=True)
            srun.advance(1, True)
            sms = srun.elapsed_ms(reset=True)
            sv = 1184 * 10 / (sms * 1e-3)
            srun.close()
            also = {"synthetic_1024x4096": {"value": sv, "unit": UNIT, "chains": 1184, "sweeps_per_chain": 10, "kernel_ms": sms,
                                             "fp64_frac": 10.0 * ALGO["synthetic"]["F"] * sv / 1e12 / mb["fp64_tflops"],
                                             "note": "sweep kernel only (ser_sweep_kernel_big); full line: --dataset synthetic"}}
        h2d = int(X.shape[0] * ((M + 31) // 32) * 4 + N + 4 * M)
        d2h = int(n_local * 200 + k * N * N * 4)
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": ("synthetic 1024 x 4096 occurrence matrix (ser_dataset_synthetic), Philox free-running chains" if big else
                                     "NOW subset %s (committed fixture of the reference's Dataset/), Philox free-running chains" % args.dataset),
            "config": {"workload": "%s %dx%d, %d chains per GPU x (%d burn + %d sampling) calls x 10 sweeps, thin 10, "
                                   "on-device selection (k=%d) + pair-order counts" % (args.dataset, N, M, n_local, args.burn_calls, args.sample_calls, k),
                       "chains_total": n_total, "sweeps_per_step": sweeps_per_step, "parallelism": "chains sharded x%d" % world,
                       "l2": "256 MB flush buffer written between timed steps"},
            "clocks": clk, "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                                   "ms_per_step": dt_e2e / args.steps * 1e3},
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "also": also,
        }))
    if world > 1:
        dist.destroy_process_group()


def _hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0  # fallback stated in B200_PROFILING.md


if __name__ == "__main__":
    main()
