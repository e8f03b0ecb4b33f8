#!/usr/bin/env python
"""Aggregate an ncu report's per-SASS-instruction counters by CUDA source line.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep ser_sweep_kernel [top]

Joins `ncu --page source --csv` (SASS view: executed instructions + stall samples per
instruction) with `nvdisasm -g` line info of the in-tree library (needs -lineinfo)."""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "seriation-in-paleontological-data-using-mcmc_b200", "libseriation_b200.so")


def line_map(kernel):
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=td, check=True, stdout=subprocess.DEVNULL)
        cub = [f for f in os.listdir(td) if f.startswith("ser_kernels")][0]
        dis = subprocess.run(["nvdisasm", "-g", "-c", cub], cwd=td, capture_output=True, text=True).stdout
    m, cur, inside = {}, None, False
    for ln in dis.split("\n"):
        if ln.startswith(".text."):
            inside = kernel in ln
            continue
        if not inside:
            continue
        f = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if f:
            cur = (os.path.basename(f.group(1)), int(f.group(2)))
            continue
        a = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if a and cur:
            m[int(a.group(1), 16)] = (cur, a.group(2).strip())
    return m


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = next(r for r in rows if r and r[0] == "Address")
    ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    data = [r for r in rows if len(r) == len(hdr) and r[0] != "Address"]
    base = min(int(r[ia], 16) for r in data)
    lm = line_map(kernel)
    agg, tot_i, tot_s = {}, 0, 0
    for r in data:
        off = int(r[ia], 16) - base
        key = lm.get(off, (("?", 0), ""))[0]
        n, s = int(r[ii] or 0), int(r[isamp] or 0)
        a = agg.setdefault(key, [0, 0])
        a[0] += n
        a[1] += s
        tot_i += n
        tot_s += s
    src = {}
    print("total warp-instructions %d, stall samples %d" % (tot_i, tot_s))
    for key, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        f, l = key
        if f not in src:
            p = os.path.join(os.path.dirname(LIB), "csrc", f)
            src[f] = open(p).read().split("\n") if os.path.exists(p) else []
        text = src[f][l - 1].strip()[:80] if 0 < l <= len(src[f]) else ""
        print("%5.1f%% inst %5.1f%% stall  %s:%d  %s" % (100.0 * n / tot_i, 100.0 * s / max(1, tot_s), f, l, text))


if __name__ == "__main__":
    main()
