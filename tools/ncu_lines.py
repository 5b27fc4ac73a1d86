#!/usr/bin/env python3
"""Per-source-line instruction and stall-sample totals for one profiled kernel (development aid).

    tools/ncu_lines.py <report.ncu-rep> <object-with-cubin (.o)> <mangled-kernel-substring>

Joins the SASS view of the ncu source page with nvdisasm's line table by instruction order."""
import csv, os, re, subprocess, sys, tempfile, collections
rep, obj, sub = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(sass.splitlines()))
hdr = rows[1]
iS, iN, iI = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
prof = [(r[iS].strip(), int(r[iI] or 0), int(r[iN] or 0)) for r in rows[2:] if len(r) > iI]
with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
lines, cur, active, loc = [], None, False, ("?", 0)
for l in dis.splitlines():
    if l.startswith(".text."):
        active = sub in l
        continue
    if not active:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        loc = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
    if m:
        lines.append((m.group(1).strip(), loc))
if len(lines) != len(prof):
    print("warning: %d disassembled vs %d profiled instructions" % (len(lines), len(prof)))
agg = collections.defaultdict(lambda: [0, 0])
tot_i = sum(p[1] for p in prof); tot_s = sum(p[2] for p in prof)
for (txt, loc), (ptxt, inst, samp) in zip(lines, prof):
    agg[loc][0] += inst; agg[loc][1] += samp
src_cache = {}
def src(loc):
    f, n = loc
    for base in ("hipims_ocl_b200/csrc",):
        p = os.path.join(base, f)
        if os.path.exists(p):
            if p not in src_cache: src_cache[p] = open(p).read().splitlines()
            return src_cache[p][n - 1].strip()[:110] if 0 < n <= len(src_cache[p]) else ""
    return ""
print("total warp instructions %d, samples %d" % (tot_i, tot_s))
for loc, (inst, samp) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.2f%% inst %5.2f%% stall  %s:%d  %s" % (100.0 * inst / tot_i, 100.0 * samp / max(1, tot_s), loc[0], loc[1], src(loc)))
