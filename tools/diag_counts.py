import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hipims_ocl_b200 import executor as hx
from tests.helpers import dtype_of, make_cfg, scenario
ex = hx.Executor(0)
for scheme in ("muscl-hancock", "godunov"):
    for opt in (0, hx.OPT_NO_GRAPH, hx.OPT_STRICT_FP):
        for rep in range(4):
            cfg = make_cfg(scheme, "double", 96, 96)
            bed, st, man = scenario("dambreak", 96, 96, np.float64)
            g = hx.CudaScheme(ex, cfg, options=opt)
            g.upload(st, bed, man); g.set_target(1e6)
            g.iterate(200)
            s = g.stats()
            print(scheme, opt, rep, s["batch_successful"], s["batch_skipped"], "%.9f %.9f" % (s["time"], s["timestep"]), g.raw_stats().iterations, g.raw_stats().kernel_launches)
            g.close()
