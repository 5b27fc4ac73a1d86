#!/usr/bin/env python3
"""BASELINE configs[0] end to end: the reference's Newcastle test case (342 x 195 cells at 2 m, rain 70 mm/h for an hour
then dry, losses 12 mm/h, Godunov fp64, 7200 s) on the CUDA executor against the CPU oracle, compared on the rasters the
reference writes (depth, maxdepth).  The DEM comes from tests/golden/newcastle_centre.npz.

    python tools/newcastle_full.py [end_time_s] [queue] [HP_OPT_* mask, 1 = strict flavour]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hipims_ocl_b200 import config as hc, executor as hx   # noqa: E402
from oracle import cpu_sim, raster_oracle as ro            # noqa: E402
from tests.test_golden import newcastle_sim                # noqa: E402

end = float(sys.argv[1]) if len(sys.argv) > 1 else 7200.0
queue = int(sys.argv[2]) if len(sys.argv) > 2 else 512
options = int(sys.argv[3]) if len(sys.argv) > 3 else 0
ex = hx.Executor(0)
runs = {}
for name, make in (("cuda", lambda cfg: hx.CudaScheme(ex, cfg.with_(end_time=end), options=options)), ("oracle", lambda cfg: cpu_sim.CpuSim("oracle", cfg.with_(end_time=end), threads=os.cpu_count()))):
    z, bed, sim = newcastle_sim(make)
    sim.set_target(end)
    t0 = time.perf_counter()
    iters = 0
    while sim.stats()["time"] < end - 1e-5:
        sim.iterate(queue)
        iters += queue
    wall = time.perf_counter() - t0
    runs[name] = (sim.download(), sim.stats(), wall, iters)
    print("%-6s t = %.3f s, %d successful iterations (%d scheduled), %.1f s wall, %.1f M cell-updates/s" % (
        name, runs[name][1]["time"], runs[name][1]["batch_successful"], iters, wall, bed.size * iters / wall / 1e6))
g, o = runs["cuda"][0], runs["oracle"][0]
for label, code in (("depth", ro.DEPTH), ("maxdepth", ro.MAX_DEPTH)):
    rg, rw = ro.derive_raster(code, g, bed, 2.0), ro.derive_raster(code, o, bed, 2.0)
    both = (rg != -9999.0) & (rw != -9999.0)
    print("%-8s raster: wet cells %d vs %d (both %d); max |diff| %.3e m, mean |diff| %.3e m, 99.9th percentile %.3e m; max value %.3f m" % (
        label, int((rg != -9999.0).sum()), int((rw != -9999.0).sum()), int(both.sum()), np.abs(rg - rw)[both].max(),
        np.abs(rg - rw)[both].mean(), np.quantile(np.abs(rg - rw)[both], 0.999), rw[both].max()))
vg, vw = (g[..., 0] - bed).sum() * 4.0, (o[..., 0] - bed).sum() * 4.0
print("volume %.3f vs %.3f m3 (relative difference %.2e); successful iterations %d vs %d" % (
    vg, vw, abs(vg - vw) / vw, runs["cuda"][1]["batch_successful"], runs["oracle"][1]["batch_successful"]))
