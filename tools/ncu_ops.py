#!/usr/bin/env python3
"""SASS opcode histogram (executed warp instructions per 32 cells) of one profiled kernel.

    tools/ncu_ops.py <report.ncu-rep> <cells>"""
import csv, collections, subprocess, sys
rep, cells = sys.argv[1], float(sys.argv[2]) / 32
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iS, iI = hdr.index("Source"), hdr.index("Instructions Executed")
ops, tot = collections.Counter(), 0
for r in rows[2:]:
    if len(r) <= iI: continue
    s = r[iS].strip(); n = int(r[iI] or 0)
    op = s.split()[1] if s.startswith("@") else s.split()[0]
    ops[op.split(".")[0]] += n; tot += n
for k, v in ops.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 30):
    print("%-12s %7.1f per 32 cells  %5.2f%%" % (k, v / cells, 100 * v / tot))
print("total %.1f, SASS lines %d" % (tot / cells, len(rows) - 2))
