"""Development aid: run a few iterations of one workload (used under ncu)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from hipims_ocl_b200 import executor as hx

name = sys.argv[1] if len(sys.argv) > 1 else "dambreak4096"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
options = int(sys.argv[3]) if len(sys.argv) > 3 else hx.OPT_NO_GRAPH
n = int(sys.argv[4]) if len(sys.argv) > 4 and int(sys.argv[4]) > 0 else None
w = dict(bench.WORKLOADS[name])
if n:
    w["cols"] = w["rows_per_gpu"] = n
cfg = bench.cfg_for(w, w["rows_per_gpu"], w["cols"])
dtype = np.float64 if cfg.precision == "double" else np.float32
bed, st, man = bench.make_inputs(w, cfg.rows, cfg.cols, dtype)
ex = hx.Executor(0)
sim = hx.CudaScheme(ex, cfg, options=options)
sim.upload(st, bed, man)
bench.attach_boundaries(sim, w, cfg.cols, cfg.rows)
sim.set_target(1e7)
if len(sys.argv) > 5:                       # spin-up: iterate until this simulated time (rain has fallen), then `iters` more
    while sim.stats()["time"] < float(sys.argv[5]):
        sim.iterate(16)
    sim.sync()
    import torch                            # ncu --profile-from-start off: only what follows is profiled
    torch.cuda.cudart().cudaProfilerStart()
sim.iterate(iters)
print(sim.stats())
