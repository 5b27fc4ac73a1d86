"""Scratch diagnostic: clock after suspension + resume, oracle vs CUDA, odd/even iteration counts, with/without the dt0 quirk."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hipims_ocl_b200 import config as hc, executor as hx
from oracle import cpu_sim
from tests.helpers import make_cfg, scenario

ex = hx.Executor(0)
n = 96
bed, st, man = scenario("dambreak", n, n, np.float64)
for quirk in (0, hc.QUIRK_GODUNOV_DT0_KEEP):
    for first in (32, 33):
        for options in (hx.OPT_STRICT_FP, hx.OPT_STRICT_FP | hx.OPT_NO_GRAPH):
            cfg = make_cfg("godunov", "double", n, n, friction=False)
            cfg.quirks |= quirk
            orc = cpu_sim.CpuSim("oracle", cfg)
            gpu = hx.CudaScheme(ex, cfg, options=options)
            for sim in (orc, gpu):
                sim.upload(st, bed, man); sim.set_target(0.25); sim.iterate(first)
            print("quirk", quirk, "first", first, "opt", options)
            print("  susp  orc", {k: orc.stats()[k] for k in ("time", "timestep", "batch_successful", "batch_skipped", "use_alternate")})
            print("  susp  gpu", {k: gpu.stats()[k] for k in ("time", "timestep", "batch_successful", "batch_skipped", "use_alternate")})
            for sim in (orc, gpu):
                sim.set_target(0.5); sim.update_timestep(); sim.reset_counters()
            print("  upd   orc", orc.stats()["timestep"], " gpu", gpu.stats()["timestep"])
            for i in range(4):
                for sim in (orc, gpu):
                    sim.iterate(1)
                so, sg = orc.stats(), gpu.stats()
                print("  it%d   orc dt=%.6g ok=%d skip=%d | gpu dt=%.6g ok=%d skip=%d" % (i, so["timestep"], so["batch_successful"], so["batch_skipped"],
                                                                                     sg["timestep"], sg["batch_successful"], sg["batch_skipped"]))
            gpu.close(); orc.close()
