"""Shared helpers for the parity tests."""
import numpy as np

from hipims_ocl_b200 import config as hc
from hipims_ocl_b200 import scenarios as sc


def dtype_of(precision):
    return np.float64 if precision == "double" else np.float32


def make_cfg(scheme, precision, rows, cols, **kw):
    base = dict(scheme=scheme, precision=precision, rows=rows, cols=cols, delta=1.0, end_time=1.0e6)
    base.update(kw)
    return hc.SchemeConfig(**base)


def add_standard_boundaries(sim, cfg, kind):
    """Attach one of the named boundary sets to any object with the add_* interface
    (CpuSim and the CUDA binding share it)."""
    if kind == "none":
        return
    if kind in ("rain", "rain+loss"):
        sim.add_uniform(hc.UNIFORM_RAIN_INTENSITY, [0.0, 3600.0, 7200.0, 10800.0], [70.0, 70.0, 0.0, 0.0])
        if kind == "rain+loss":
            sim.add_uniform(hc.UNIFORM_LOSS_RATE, [0.0, 1.0e8], [12.0, 12.0])
        return
    if kind == "gridded":
        rng = np.random.default_rng(7)
        res = 8.0
        gr, gc = int(np.ceil(cfg.rows * cfg.delta / res)) + 1, int(np.ceil(cfg.cols * cfg.delta / res)) + 1
        frames = rng.uniform(0.0, 80.0, size=(5, gr, gc))
        sim.add_gridded(hc.GRIDDED_RAIN_INTENSITY, 300.0, res, 0.0, 0.0, frames)
        return
    if kind == "cells":
        rows, cols = cfg.rows, cfg.cols
        # imposed discharge along part of the west edge (dischargeValue=total -> host divides by count)
        west = [y * cols + 1 for y in range(rows // 3, 2 * rows // 3)]
        ts = np.array([[0.0, 0.0, 0.0, 0.0], [60.0, 0.0, 40.0, 0.0], [1.0e5, 0.0, 40.0, 0.0]])
        ts[:, 2:] /= len(west)
        sim.add_cell(hc.DEPTH_IGNORE, hc.DISCHARGE_IS_DISCHARGE, west, ts)
        # imposed free-surface level on the east edge
        east = [y * cols + cols - 2 for y in range(rows // 4, 3 * rows // 4)]
        tide = np.array([[0.0, 1.5, 0.0, 0.0], [100.0, 2.5, 0.0, 0.0], [1.0e5, 2.5, 0.0, 0.0]])
        sim.add_cell(hc.DEPTH_IS_FSL, hc.DISCHARGE_IGNORE, east, tide)
        # surcharging sewers: point volume sources
        pts = [(rows // 2) * cols + cols // 2, (rows // 3) * cols + cols // 3]
        vol = np.array([[0.0, 0.0, 0.0, 0.0], [30.0, 0.0, 2.0, 0.0], [1.0e5, 0.0, 0.5, 0.0]])
        sim.add_cell(hc.DEPTH_IGNORE, hc.DISCHARGE_IS_VOLUME, pts, vol)
        # imposed depth + velocity
        vel = [(rows // 5) * cols + cols // 5]
        vts = np.array([[0.0, 0.4, 0.3, -0.2], [1.0e5, 0.6, 0.3, -0.2]])
        sim.add_cell(hc.DEPTH_IS_DEPTH, hc.DISCHARGE_IS_VELOCITY, vel, vts)
        return
    raise ValueError(kind)


def scenario(name, rows, cols, dtype, seed=11):
    if name == "dambreak":
        assert rows == cols
        return sc.dam_break(rows, dtype=dtype)
    if name == "dambreak-dry":
        return sc.dam_break(rows, outer_level=0.0, dtype=dtype)
    if name == "wetdry":
        return sc.random_wet_dry(rows, cols, seed, dtype=dtype)
    if name == "pluvial":
        return sc.pluvial(rows, cols, seed=seed, dtype=dtype)
    if name == "pluvial-wet":
        return sc.pluvial(rows, cols, seed=seed, dtype=dtype, wet_fraction=0.3)
    if name == "valley":
        return sc.river_valley(rows, cols, seed=seed, dtype=dtype)
    if name == "lake":
        return sc.lake_at_rest(rows, cols, seed=seed, dtype=dtype)
    raise ValueError(name)
