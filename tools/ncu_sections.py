#!/usr/bin/env python3
"""Issue-cycle budget of one profiled kernel by OUTERMOST source line (inlined helpers are charged to their call site).

    tools/ncu_sections.py <report.ncu-rep> <object-with-cubin (.o)> <mangled-kernel-substring> <kernel-file-basename> <units>

`units` = what to normalise by (e.g. warp rows = cells / 30).  Cost model of profiles/r02_fp64_issue_probe3.txt: an fp64
instruction 2.23 cycles, an IMAD 2.0, everything else 1.1."""
import csv, os, re, subprocess, sys, tempfile, collections
rep, obj, sub, kfile, units = sys.argv[1:6]
units = float(units)
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(sass.splitlines()))
hdr = rows[1]
iS, iI = hdr.index("Source"), hdr.index("Instructions Executed")
prof = [(r[iS].strip(), int(r[iI] or 0)) for r in rows[2:] if len(r) > iI]
with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
lines, active, chain = [], False, []
pending = []
for l in dis.splitlines():
    if l.startswith("//-") and ".text." in l:
        active = sub in l
        continue
    if not active:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        pending.append((os.path.basename(m.group(1)), int(m.group(2))))
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
    if m:
        if pending:
            chain = pending; pending = []
        outer = [c for c in chain if c[0] == kfile]
        lines.append((m.group(1).strip(), outer[-1][1] if outer else -1))
if len(lines) != len(prof):
    print("warning: %d disassembled vs %d profiled instructions" % (len(lines), len(prof)))
def cost(txt):
    t = txt.split()
    op = t[1] if txt.startswith("@") else t[0]
    b = op.split(".")[0]
    if b in ("DADD", "DMUL", "DFMA", "DSETP"): return 2.23, "f64"
    if b == "IMAD": return 2.0, "imad"
    return 1.1, "other"
agg = collections.defaultdict(lambda: collections.Counter())
for (txt, line), (ptxt, inst) in zip(lines, prof):
    c, cls = cost(txt)
    agg[line]["cyc"] += c * inst; agg[line][cls] += inst; agg[line]["n"] += inst
tot = sum(v["cyc"] for v in agg.values())
src = open(os.path.join(sys.argv[6] if len(sys.argv) > 6 else "hipims_ocl_b200/csrc", kfile)).read().splitlines()
print("model cycles per unit: %.0f   instructions per unit: %.0f" % (tot / units, sum(v["n"] for v in agg.values()) / units))
print(" line   cyc/unit   %%    n    f64  imad  other")
for line in sorted(agg):
    v = agg[line]
    if v["cyc"] / units < 1.0: continue
    print("%5d  %7.1f  %4.1f  %5.1f %5.1f %5.1f %5.1f  %s" % (line, v["cyc"] / units, 100 * v["cyc"] / tot, v["n"] / units, v["f64"] / units,
                                                       v["imad"] / units, v["other"] / units, src[line - 1].strip()[:90] if 0 < line <= len(src) else ""))
