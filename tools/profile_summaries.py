#!/usr/bin/env python3
"""Turns the captures of tools/profile_round.sh into the text summaries and kernels.json kept under profiles/.

    python tools/profile_summaries.py <tag> [<round label>]          (needs ncu to read the reports; no GPU)
"""
import csv, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
label = sys.argv[2] if len(sys.argv) > 2 else tag
CELLS = {"dambreak4096": 4096 * 4096, "dambreak4096-f32": 4096 * 4096, "dambreak4096-mh": 4096 * 4096, "dambreak4096-mh-f32": 4096 * 4096,
         "dambreak4096-inertial": 4096 * 4096, "dambreak4096-inertial-f32": 4096 * 4096, "pluvial16384": 16384 * 16384,
         "river32768": 32768 * 4096}
NOTE = {"pluvial16384": "configs[2], 16384 x 16384, after the rain has fallen twice (a film of water on every cell)",
        "river32768": "configs[4], one 32768 x 4096 strip (river along the valley floor, dry valley sides)"}
FP64 = ("DFMA", "DADD", "DMUL", "DSETP")
out = {"_comment": "per-launch figures of the dominant kernel of each workload from the ncu --set full captures summarised in this "
                   "directory (%s; one launch after warm-up / spin-up); bench.py scales dram_bytes_per_cell by the cells one launch "
                   "covers for roofline.traffic and uses the instruction counts for roofline_issue" % label}
for wl, cells in CELLS.items():
    rep = os.path.join(ROOT, "gpurun_out", "%s_%s.ncu-rep" % (tag, wl))
    if not os.path.exists(rep):
        continue
    txt = os.path.join(ROOT, "profiles", "%s_%s.txt" % (label, wl.replace("-", "_")))
    head = "ncu --set full --clock-control none --import-source on, one launch after warm-up (tools/run_short.py); B200, %s" % \
           NOTE.get(wl, "4096 x 4096 dam break")
    res = subprocess.run(["bash", os.path.join(ROOT, "tools", "ncu_profile_txt.sh"), rep, str(cells), head], capture_output=True, text=True, cwd=ROOT)
    open(txt, "w").write(res.stdout)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]
    def val(k):
        return float(r[hdr.index(k)].replace(",", ""))
    def scale(k):
        return {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3, "us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}[units[hdr.index(k)]]
    dram = (val("dram__bytes_read.sum") + val("dram__bytes_write.sum")) * scale("dram__bytes_read.sum")
    ops = {}
    for line in res.stdout.splitlines():
        m = re.match(r"^([A-Z0-9_]+)\s+([0-9.]+) per 32 cells", line)
        if m:
            ops[m.group(1)] = float(m.group(2))
    inst = val("smsp__inst_executed.sum") / (cells / 32.0)
    out[wl] = {"kernel": r[hdr.index("Kernel Name")].split("(")[0], "profile": os.path.basename(txt),
               "dram_bytes_per_launch": dram, "dram_bytes_per_cell": dram / cells,
               "ncu_duration_ms": val("gpu__time_duration.sum") * scale("gpu__time_duration.sum"),
               "warp_instructions_per_32_cells": round(inst, 1),
               "fp64_instructions_per_32_cells": round(sum(ops.get(k, 0.0) for k in FP64), 1),
               "fp64_pipe_pct": round(val("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"), 1),
               "alu_pipe_pct": round(val("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"), 1),
               "issue_active_pct": round(val("smsp__issue_active.avg.pct_of_peak_sustained_active"), 1),
               "dram_throughput_pct": round(val("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), 1),
               "registers": int(val("launch__registers_per_thread"))}
    print(wl, out[wl])
json.dump(out, open(os.path.join(ROOT, "profiles", "kernels.json"), "w"), indent=1)
launches = os.path.join(ROOT, "gpurun_out", "%s_launches.csv" % tag)
if os.path.exists(launches):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_launches.py"), launches,
                          "python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-variants"], capture_output=True, text=True)
    open(os.path.join(ROOT, "profiles", "%s_launches_bench.txt" % label), "w").write(res.stdout.replace("round 1", label))
    print(res.stdout[:1500])
