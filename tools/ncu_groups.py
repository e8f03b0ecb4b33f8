#!/usr/bin/env python
"""Group tools/ncu_lines.py output into code regions (by function) for the sweep kernel."""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CS = os.path.join(ROOT, "seriation-in-paleontological-data-using-mcmc_b200", "csrc")

def func_ranges(path):
    """(start_line, name) for every top-level function-ish definition"""
    out = []
    for i, l in enumerate(open(path).read().split("\n"), 1):
        m = re.match(r"^(?:template <[^>]*>\s*)?(?:SER_HD|__device__|__global__|static|extern \"C\"|__host__)[^;{]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", l)
        if m and not l.startswith(" "):
            out.append((i, m.group(1)))
    return out

def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, kernel, "100000"], capture_output=True, text=True).stdout.split("\n")
    print(txt[0])
    ranges = {f: func_ranges(os.path.join(CS, f)) for f in ("ser_chain_core.h", "ser_kernels.cu", "ser_detmath.h", "ser_device_common.cuh", "ser_sweep_kernel.cuh",
                                                                     "ser_sweep_kernel_big.cuh", "ser_aux_kernels.cuh")}
    agg = {}
    for ln in txt[1:]:
        m = re.match(r"\s*([\d.]+)% inst\s+([\d.]+)% stall\s+(\S+):(\d+)", ln)
        if not m: continue
        f, l = m.group(3), int(m.group(4))
        name = f
        if f in ranges:
            cand = [n for (s, n) in ranges[f] if s <= l]
            name = f.split(".")[0][4:] + ":" + (cand[-1] if cand else "?")
        a = agg.setdefault(name, [0.0, 0.0]); a[0] += float(m.group(1)); a[1] += float(m.group(2))
    for n, (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
        print("%6.1f%% inst %6.1f%% stall  %s" % (i, s, n))

if __name__ == "__main__":
    main()
