#!/usr/bin/env python3
"""Development aid: growth of the CUDA-vs-oracle difference over iterations for one case.

    python tools/diag_parity.py SCHEME PRECISION SCENARIO BOUNDARIES N ITERS [key=value ...]
"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hipims_ocl_b200 import executor as hx
from oracle import cpu_sim
from tests.helpers import add_standard_boundaries, dtype_of, make_cfg, scenario


def main():
    scheme, precision, scen, bdy, n, iters = sys.argv[1:7]
    n, iters = int(n), int(iters)
    extra = {}
    for kv in sys.argv[7:]:
        k, v = kv.split("=")
        extra[k] = eval(v)
    every = extra.pop("every", 20)
    cfg = make_cfg(scheme, precision, n, n, **extra)
    bed, st, man = scenario(scen, n, n, dtype_of(precision))
    ex = hx.Executor(0)
    sims = {"oracle": cpu_sim.CpuSim("oracle", cfg), "strict": hx.CudaScheme(ex, cfg, options=hx.OPT_STRICT_FP),
            "fast": hx.CudaScheme(ex, cfg, options=0)}
    for s in sims.values():
        s.upload(st, bed, man)
        add_standard_boundaries(s, cfg, bdy)
        s.set_target(1e6)
    done = 0
    while done < iters:
        for s in sims.values():
            s.iterate(every)
        done += every
        ref = sims["oracle"].download()
        row = ["it %4d t=%.4f" % (done, sims["oracle"].stats()["time"])]
        for name in ("strict", "fast"):
            cur = sims[name].download()
            d = np.abs(cur[..., 0].astype(np.float64) - ref[..., 0])
            i = np.unravel_index(np.argmax(d), d.shape)
            dq = np.abs(cur[..., 2:].astype(np.float64) - ref[..., 2:]).max()
            row.append("%s: dEta=%.3e at %s (h=%.3e) dQ=%.3e dt_diff=%.2e" % (
                name, d.max(), i, ref[i][0] - bed[i], dq, sims[name].stats()["timestep"] - sims["oracle"].stats()["timestep"]))
        print(" | ".join(row))


if __name__ == "__main__":
    main()
