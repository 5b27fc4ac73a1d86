"""Development aid: device-resident throughput of several workloads, one line each."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from hipims_ocl_b200 import executor as hx

names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["dambreak4096", "dambreak4096-f32", "dambreak4096-mh"]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
options = int(sys.argv[3]) if len(sys.argv) > 3 else 0
ex = hx.Executor(0)
peak, _ = bench.peak_hbm()
for name in names:
    spin = None
    if "@" in name:                      # name[:size]@T -- spin up until simulated time T (rain has fallen) before timing
        name, spin = name.split("@")[0], float(name.split("@")[1])
    w = dict(bench.WORKLOADS[name.split(":")[0]])
    if ":" in name:
        w["cols"] = w["rows_per_gpu"] = int(name.split(":")[1])
    cfg = bench.cfg_for(w, w["rows_per_gpu"], w["cols"])
    dtype = np.float64 if cfg.precision == "double" else np.float32
    bed, st, man = bench.make_inputs(w, cfg.rows, cfg.cols, dtype)
    sim = hx.CudaScheme(ex, cfg, options=options)
    sim.upload(st, bed, man)
    bench.attach_boundaries(sim, w, cfg.cols, cfg.rows)
    sim.set_target(1e7)
    sim.iterate(10)
    while spin is not None and sim.stats()["time"] < spin:
        sim.iterate(32)
    best = 0.0
    for rep in range(3):
        ex.timer_start()
        sim.iterate(steps, sync=False)
        ms = ex.timer_stop()
        best = max(best, cfg.cells * steps / (ms * 1e-3))
    print("%-22s %-14s %s  %7.2f G cell-updates/s  roofline %.1f%%  (t=%.2f s)" % (
        name, cfg.scheme, cfg.precision, best / 1e9, 100 * best * bench.algorithmic_bytes_per_cell(cfg) / 1e9 / peak, sim.stats()["time"]))
    sim.close()
