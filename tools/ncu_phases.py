#!/usr/bin/env python3
"""Splits the SASS of a profiled kernel at BAR.SYNC / loop markers and prints instruction and
stall-sample totals per segment (development aid)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iN, iI = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
seg, segs = {"name": "start", "inst": 0, "samp": 0, "n": 0, "ops": {}}, []
tot_i = tot_s = 0
for r in rows[2:]:
    if len(r) <= iI: continue
    src = r[iS].strip()
    inst, samp = int(r[iI] or 0), int(r[iN] or 0)
    seg["inst"] += inst; seg["samp"] += samp; seg["n"] += 1
    op = src.split()[1] if src.startswith("@") else src.split()[0]
    op = op.split(".")[0]
    seg["ops"][op] = seg["ops"].get(op, 0) + inst
    tot_i += inst; tot_s += samp
    if src.startswith("BAR.") or " BAR." in src or "SYNCS.PHASECHK" in src or "UTMALDG" in src:
        segs.append(seg)
        seg = {"name": src[:40], "inst": 0, "samp": 0, "n": 0, "ops": {}}
segs.append(seg)
print("total warp instructions %d, samples %d" % (tot_i, tot_s))
for s in segs:
    if s["inst"] == 0 and s["samp"] == 0: continue
    top = sorted(s["ops"].items(), key=lambda kv: -kv[1])[:8]
    print("after %-42s sass=%4d inst=%6.2f%% samples=%6.2f%%  %s" % (s["name"], s["n"], 100.0 * s["inst"] / tot_i, 100.0 * s["samp"] / max(1, tot_s),
          " ".join("%s:%.1f%%" % (k, 100.0 * v / tot_i) for k, v in top)))
