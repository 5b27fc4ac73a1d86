#!/usr/bin/env python3
"""Development aid: growth of the CUDA-vs-oracle difference on the Newcastle case (configs[0])."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hipims_ocl_b200 import executor as hx
from oracle import cpu_sim
from tests.test_golden import newcastle_sim

every = int(sys.argv[1]) if len(sys.argv) > 1 else 10
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 120
ex = hx.Executor(0)
sims = {}
z, bed, sims["oracle"] = newcastle_sim(lambda cfg: cpu_sim.CpuSim("oracle", cfg))
_, _, sims["strict"] = newcastle_sim(lambda cfg: hx.CudaScheme(ex, cfg, options=hx.OPT_STRICT_FP))
_, _, sims["fast"] = newcastle_sim(lambda cfg: hx.CudaScheme(ex, cfg, options=0))
done = 0
while done < iters:
    for s in sims.values():
        s.iterate(every)
    done += every
    ref = sims["oracle"].download()
    row = ["it %4d t=%.4f wet=%d" % (done, sims["oracle"].stats()["time"], int(((ref[..., 0] - bed) > 1e-10).sum()))]
    for name in ("strict", "fast"):
        cur = sims[name].download()
        d = np.abs(cur[..., 0] - ref[..., 0])
        i = np.unravel_index(np.argmax(d), d.shape)
        row.append("%s: dEta=%.3e at %s (h=%.3e) n>1e-12=%d dQ=%.3e" % (name, d.max(), i, ref[i][0] - bed[i], int((d > 1e-12).sum()),
                                                                   np.abs(cur[..., 2:] - ref[..., 2:]).max()))
    print(" | ".join(row))
