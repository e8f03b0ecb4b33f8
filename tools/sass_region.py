"""static SASS view of one kernel: python tools/sass_region.py <nvdisasm -g dump of one function> [first_line last_line of <file>]
Prints the instructions with their source line; with a range, only the instructions attributed to
ser_sweep_kernel.cuh lines in that range (and inlined callees between them)."""
import re
import sys

path = sys.argv[1]
cur_file, cur_line = None, None
rows = []
for ln in open(path):
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_file, cur_line = m.group(1).split("/")[-1], int(m.group(2))
        continue
    m = re.match(r'\s*/\*([0-9a-f]+)\*/\s+(.*?);', ln)
    if m:
        rows.append((int(m.group(1), 16), cur_file, cur_line, m.group(2).strip()))
    elif re.match(r'\s*\.L_x_\d+:', ln):
        rows.append((-1, None, None, ln.strip()))
if len(sys.argv) >= 4:
    lo, hi = int(sys.argv[2]), int(sys.argv[3])
    inside = False
    for addr, f, l, ins in rows:
        if f == "ser_sweep_kernel.cuh" and l is not None:
            inside = lo <= l <= hi
        if inside or addr < 0:
            print("%6s %-26s %s" % ("%x" % addr if addr >= 0 else "", "%s:%s" % (f, l) if f else "", ins))
else:
    for addr, f, l, ins in rows:
        print("%6s %-26s %s" % ("%x" % addr if addr >= 0 else "", "%s:%s" % (f, l) if f else "", ins))
