#!/usr/bin/env python3
"""Development aid: N row strips stepped one after the other on ONE GPU, halo rows moved by the host through
hp_scheme_read_rows / write_rows after every iteration (the reference's own CDomainLink protocol), fixed timestep --
against the whole domain in one scheme.  Checks that the kernels do not depend on the decomposition.

    python tools/strips_emulated.py [scheme] [precision] [rows] [cols] [strips] [iters] [bdy] [options]
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hipims_ocl_b200 import executor as hx, strips
from tests.helpers import add_standard_boundaries, dtype_of, make_cfg, scenario


def run(scheme="muscl-hancock", precision="double", rows=1024, cols=512, n=8, iters=40, bdy="cells", options=0, verbose=True):
    dt = dtype_of(precision)
    bed, st, man = scenario("valley", rows, cols, dt)
    cfg_full = make_cfg(scheme, precision, rows, cols, dynamic=False, fixed_dt=0.02)
    ex = hx.Executor(0)
    ref = hx.CudaScheme(ex, cfg_full, options=options)
    ref.upload(st, bed, man)
    add_standard_boundaries(ref, cfg_full, bdy)
    ref.set_target(1e6)
    parts = []
    for r in range(n):
        s = strips.make_strip(rows, n, r, scheme)
        sim = hx.CudaScheme(ex, cfg_full.with_(rows=s.rows), options=options | hx.OPT_NO_GRAPH, global_rows=rows, row_offset=s.row_offset,
                            halo_south=s.halo_south, halo_north=s.halo_north)
        sl = s.local_slice()
        sim.upload(st[sl], bed[sl], man[sl])
        add_standard_boundaries(sim, cfg_full, bdy)
        sim.set_target(1e6)
        parts.append((s, sim))
    halo = strips.halo_rows(scheme)
    worst = 0.0
    for it in range(iters):
        ref.iterate(1)
        for s, sim in parts:
            sim.iterate(1)
        for i in range(n - 1):                      # exchange: my top owned rows -> northern neighbour's southern halo, and back
            (sa, a), (sb, b) = parts[i], parts[i + 1]
            top = a.read_rows(sa.halo_south + sa.own_rows - halo, halo)
            bot = b.read_rows(sb.halo_south, halo)
            b.write_rows(0, top)
            a.write_rows(sa.halo_south + sa.own_rows, bot)
        want = ref.download()
        got = np.concatenate([sim.download()[s.owned_local_slice()] for s, sim in parts], axis=0)
        d = np.abs(got - want)
        if d.max() > 0 and verbose:
            y, x, c = np.unravel_index(np.argmax(d), d.shape)
            print("iteration %d: max diff %.3e at row %d col %d component %d (strip of %d rows: local row %d); cells differing %d" % (
                it + 1, d.max(), y, x, c, rows // n, y % (rows // n), int((d.max(axis=2) > 0).sum())))
            ys = np.unique(np.nonzero(d.max(axis=2) > 0)[0])
            print("   rows:", ys[:30])
        worst = max(worst, float(d.max()))
        if worst > 0:
            break
    for _, sim in parts:
        sim.close()
    ref.close()
    ex.close()
    return worst


if __name__ == "__main__":
    a = sys.argv[1:]
    w = run(a[0] if len(a) > 0 else "muscl-hancock", a[1] if len(a) > 1 else "double", int(a[2]) if len(a) > 2 else 1024,
            int(a[3]) if len(a) > 3 else 512, int(a[4]) if len(a) > 4 else 8, int(a[5]) if len(a) > 5 else 40,
            a[6] if len(a) > 6 else "cells", int(a[7]) if len(a) > 7 else 0)
    print("strips_emulated:", "IDENTICAL" if w == 0 else "DIFFERENT (max %.3e)" % w)
    sys.exit(0 if w == 0 else 1)
