#!/usr/bin/env python3
"""Generates tests/golden/*.npz from the REFERENCE's own kernels (oracle/_ref, built from
/root/reference by oracle/build_ref.py).  Run in the build container:

    python tests/golden/make_golden.py

Each file holds the seeded inputs, the boundary set name, and the reference outputs (both
ping-pong buffers and the clock) after a fixed number of iterations.  The fixtures travel to the
GPU box, where /root/reference does not exist.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import cpu_sim  # noqa: E402
from tests.helpers import add_standard_boundaries, dtype_of, make_cfg, scenario  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (scheme, precision, scenario, boundaries, rows, cols, iterations, extra config)
    # adversarial inputs (random rough bed, patchy water, shocks everywhere): friction off so that
    # no pow() is involved and a strict-IEEE implementation must reproduce them bit for bit
    "godunov_f64_wetdry_rain": ("godunov", "double", "wetdry", "rain", 40, 36, 30, {"friction": False}),
    "godunov_f32_wetdry_rain": ("godunov", "single", "wetdry", "rain", 40, 36, 30, {"friction": False}),
    "godunov_f64_dambreak_nofric": ("godunov", "double", "dambreak-dry", "none", 40, 40, 40, {"friction": False}),
    "godunov_f64_valley_cells": ("godunov", "double", "valley", "cells", 36, 40, 60, {}),
    "godunov_f64_pluvial_gridded": ("godunov", "double", "pluvial-wet", "gridded", 36, 40, 60, {}),
    "mh_f64_dambreak_dry": ("muscl-hancock", "double", "dambreak-dry", "none", 40, 40, 30, {}),
    "mh_f32_wetdry": ("muscl-hancock", "single", "wetdry", "none", 38, 42, 30, {"friction": False}),
    "mh_f64_valley_cells": ("muscl-hancock", "double", "valley", "cells", 36, 40, 60, {}),
    "godunov_f64_dambreak_fric": ("godunov", "double", "dambreak", "none", 40, 40, 60, {}),
    "inertial_f64_valley_cells": ("inertial", "double", "valley", "cells", 36, 40, 60, {}),
    "inertial_f32_pluvial_rain": ("inertial", "single", "pluvial-wet", "rain", 36, 40, 60, {}),
}


def newcastle():
    """BASELINE.json configs[0]: the reference's own test case (test/newcastle-centre.xml), read through the host
    mirror's XML/HFA readers, stepped by the reference's kernels: Godunov fp64, friction, rain 70 mm/h + losses
    12 mm/h, reference launch coverage (Q6: columns 336+ and rows 192+ get no rain)."""
    import ctypes as C
    from hipims_ocl_b200 import build as hpbuild, config as hc
    hpbuild.build_host()
    lib = C.CDLL(hpbuild.HOST_LIB)
    lib.hph_model_load.restype = C.c_void_p
    h = lib.hph_model_load(b"/root/reference/test/newcastle-centre.xml", 1)
    for fn in ("hph_model_states", "hph_model_bed", "hph_model_manning"):
        getattr(lib, fn).restype = C.POINTER(C.c_double); getattr(lib, fn).argtypes = [C.c_void_p]
    rows, cols = 195, 342
    st = np.ctypeslib.as_array(lib.hph_model_states(h), shape=(rows, cols, 4)).copy()
    bed = np.ctypeslib.as_array(lib.hph_model_bed(h), shape=(rows, cols)).copy()
    man = np.ctypeslib.as_array(lib.hph_model_manning(h), shape=(rows, cols)).copy()
    cfg = make_cfg("godunov", "double", rows, cols, delta=2.0, end_time=7200.0)
    sim = cpu_sim.CpuSim("ref", cfg)
    sim.upload(st, bed, man)
    sim.add_uniform(hc.UNIFORM_LOSS_RATE, [0.0, 1.0e8], [12.0, 12.0])
    sim.add_uniform(hc.UNIFORM_RAIN_INTENSITY, [0.0, 3600.0, 7200.0, 10800.0], [70.0, 70.0, 0.0, 0.0])
    sim.set_target(7200.0)
    outs = {}
    done = 0
    for it in (40, 100, 600):
        sim.iterate(it - done)
        done = it
        outs[it] = (sim.download(), sim.stats())
    assert np.array_equal(np.rint(bed * 1e4) / 1e4, bed)
    np.savez_compressed(os.path.join(HERE, "newcastle_centre.npz"), bed_e4=np.rint(bed * 1e4).astype(np.int32),
                        out_40=outs[40][0], out_100=outs[100][0], out_600=outs[600][0],
                        stats_40=np.array([outs[40][1][k] for k in sorted(outs[40][1])]),
                        stats_100=np.array([outs[100][1][k] for k in sorted(outs[100][1])]),
                        stats_600=np.array([outs[600][1][k] for k in sorted(outs[600][1])]),
                        stats_keys=np.array(sorted(outs[100][1])))
    print("wrote newcastle_centre.npz (t = %.4f s after 600 iterations, %d wet cells)" % (
        outs[600][1]["time"], int(((outs[600][0][..., 0] - bed) > 1e-10).sum())))


def raster_values():
    """Output-raster derivation: the reference's own per-cell switch (CRasterDataset::domainToRaster, compiled by
    oracle/build_ref.py:build_raster) on adversarial cells -- thresholds at 1e-8, dry, disabled, bed above 9999."""
    import ctypes as C
    from oracle import build_ref
    lib = C.CDLL(build_ref.build_raster())
    lib.ref_raster_value.restype = C.c_double
    lib.ref_raster_value.argtypes = [C.c_ubyte, C.POINTER(C.c_double), C.c_double, C.c_double]
    rng = np.random.default_rng(20260817)
    rows, cols, res = 24, 40, 2.0
    bed = np.round(rng.uniform(-5.0, 60.0, size=(rows, cols)), 4)
    depth = np.where(rng.random((rows, cols)) < 0.4, 0.0, rng.uniform(0.0, 3.0, size=(rows, cols)))
    depth.flat[::7] = rng.choice([1e-8, 0.99e-8, 1.01e-8, 1e-10, 5e-9], size=depth.flat[::7].size)     # around the 1e-8 rule
    st = np.zeros((rows, cols, 4))
    st[..., 0] = bed + depth
    st[..., 1] = st[..., 0] + np.where(rng.random((rows, cols)) < 0.5, 0.0, rng.uniform(0.0, 1.0, size=(rows, cols)))
    st[..., 2:] = rng.normal(size=(rows, cols, 2)) * (depth > 0)[..., None]
    st[3, 5, 1] = -9999.0                      # disabled cell
    bed[7, 9] = 9999.9; st[7, 9, :2] = 9999.9  # closed-edge bed (CDomainCartesian.cpp:773-799)
    bed[8, 9] = 10000.5; st[8, 9, :2] = 10001.0
    out = {}
    for code in range(12):
        a = np.empty((rows, cols))
        for y in range(rows):
            for x in range(cols):
                cell = np.ascontiguousarray(st[y, x])
                a[rows - 1 - y, x] = lib.ref_raster_value(code, cell.ctypes.data_as(C.POINTER(C.c_double)), float(bed[y, x]), res)
        out["value_%d" % code] = a
    np.savez_compressed(os.path.join(HERE, "raster_values.npz"), states=st, bed=bed, resolution=res, **out)
    print("wrote raster_values.npz (%d cells x 12 value codes from the reference's switch)" % (rows * cols))


def main():
    if os.path.exists("/root/reference/test/newcastle-centre.xml"):
        newcastle()
        raster_values()
    for name, (scheme, precision, scen, bdy, rows, cols, iters, extra) in CASES.items():
        cfg = make_cfg(scheme, precision, rows, cols, **extra)
        bed, st, man = scenario(scen, rows, cols, dtype_of(precision))
        sim = cpu_sim.CpuSim("ref", cfg)
        sim.upload(st, bed, man)
        add_standard_boundaries(sim, cfg, bdy)
        sim.set_target(1.0e6)
        sim.iterate(iters)
        a, b = sim.download_both()
        stats = sim.stats()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), bed=bed, states=st, manning=man, out_a=a, out_b=b,
                            out_current=sim.download(), stats_keys=np.array(sorted(stats)),
                            stats_vals=np.array([stats[k] for k in sorted(stats)], dtype=np.float64),
                            meta=np.array([scheme, precision, scen, bdy, str(iters), repr(extra)]))
        print("wrote %s.npz (t = %.4f s)" % (name, stats["time"]))


if __name__ == "__main__":
    main()
