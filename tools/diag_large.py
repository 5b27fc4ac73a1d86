#!/usr/bin/env python3
"""Development aid: where does a large-size parity case deviate from the oracle?  (workload rows cols iters t0)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from hipims_ocl_b200 import executor as hx
from oracle import cpu_sim

name, rows, cols, iters = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
t0 = float(sys.argv[5]) if len(sys.argv) > 5 else None
hydro = float(sys.argv[6]) if len(sys.argv) > 6 else 0.97
w = dict(bench.WORKLOADS[name]); w.update(cols=cols, rows_per_gpu=rows)
cfg = bench.cfg_for(w, rows, cols)
dtype = np.float64 if cfg.precision == "double" else np.float32
bed, st, man = bench.make_inputs(w, rows, cols, dtype)
ex = hx.Executor(0)
def run(sim):
    sim.upload(st, bed, man)
    bench.attach_boundaries(sim, w, cols, rows)
    sim.set_target(1e7)
    if t0 is not None:
        sim.set_clock(t0, cfg.initial_dt, hydro)
    out = []
    for i in range(iters):
        sim.iterate(1)
        out.append(sim.download().copy())
    return out, sim.stats()
o_out, o_st = run(cpu_sim.CpuSim("oracle", cfg))
for label, opt in (("strict", hx.OPT_STRICT_FP), ("fast", 0), ("fast-tiles", hx.OPT_TILE_KERNELS | hx.OPT_MARCH_GODUNOV), ("fast-nograph", hx.OPT_NO_GRAPH)):
    g_out, g_st = run(hx.CudaScheme(ex, cfg, options=opt))
    print(label, "dt", g_st["timestep"], o_st["timestep"])
    for i in range(iters):
        d = np.abs(g_out[i][..., 0] - o_out[i][..., 0])
        y, x = np.unravel_index(np.argmax(d), d.shape)
        h = o_out[i][y, x, 0] - bed[y, x]
        print("  it %2d max|d eta| %.3e at (y=%d, x=%d) h=%.3e  n(>1e-9)=%d  n(>1e-12)=%d  dq %.3e" % (
            i + 1, d.max(), y, x, h, int((d > 1e-9).sum()), int((d > 1e-12).sum()), np.abs(g_out[i][..., 2:] - o_out[i][..., 2:]).max()))
    if label == "fast":
        d = np.abs(g_out[-1][..., 0] - o_out[-1][..., 0])
        ys, xs = np.nonzero(d > 1e-9)
        print("  rows with deviations:", np.unique(ys)[:40], "cols:", np.unique(xs)[:40])
