#!/usr/bin/env python
"""Executed warp-instructions and stall samples of an ncu capture per REGION of a kernel's own source file.

    python tools/ncu_phases.py <file.ncu-rep> <kernel substring> <kernel source file> <line>=<name> ...

Every SASS instruction is attributed, in address order, to the last instruction before it whose line info points into the
kernel's own file (inlined helpers sit next to their call site), and that line to the region starting at or before it."""
import sys
sys.path.insert(0, __import__("os").path.dirname(__file__))
import csv, io, subprocess
from ncu_lines import line_map


def main():
    rep, kernel, own = sys.argv[1], sys.argv[2], sys.argv[3]
    regions = sorted((int(a.split("=")[0]), a.split("=")[1]) for a in sys.argv[4:])
    lm = line_map(kernel)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(r for r in rows if r and r[0] == "Address")
    ca, ci, cs, ct = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
    data = [r for r in rows if len(r) == len(hdr) and r[0] != "Address"]
    base = min(int(r[ca], 16) for r in data)
    recs = sorted((int(r[ca], 16) - base, int(r[ci] or 0), int(r[cs] or 0), int(r[ct] or 0)) for r in data)
    agg, cur = {}, None
    for off, n, s, t in recs:
        (f, l), _ = lm.get(off, ((None, 0), ""))
        if f == own:
            cur = l
        name = "?"
        if cur is not None:
            for start, nm in regions:
                if start <= cur:
                    name = nm
        a = agg.setdefault(name, [0, 0, 0])
        a[0] += n; a[1] += s; a[2] += t
    tot = [sum(a[i] for a in agg.values()) for i in range(3)]
    print("total warp-instructions %d, stall samples %d" % (tot[0], tot[1]))
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print("%6.1f%% inst %6.1f%% stall  lanes %5.1f  %s" % (100.0 * a[0] / max(tot[0], 1), 100.0 * a[1] / max(tot[1], 1), a[2] / max(a[0], 1), name))


if __name__ == "__main__":
    main()
