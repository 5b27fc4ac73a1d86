"""Synthetic inputs of the shapes BASELINE.json names.

Everything here is host-side numpy: bed elevation, initial state in the reference's host
layout (cell state = {eta, eta_max, qx, qy}, row 0 = south; src/Domain/CDomain.h:28-33) and
Manning's n.  Input values are rounded to four decimals the way the reference ingests
rasters (src/util.cpp:79-93: negative values round towards -infinity).
"""
import numpy as np


def round4(a):
    """Util::round(value, 4) of the reference (src/util.cpp:79-93)."""
    m = np.asarray(a, dtype=np.float64) * 10000
    rem = np.fmod(m, 1)                       # negative for negative values, so they always floor
    return np.where(rem >= 0.5, np.ceil(m), np.floor(m)) / 10000


def make_states(bed, depth, qx=None, qy=None, dtype=np.float64):
    """Initial cell states: eta = bed + depth, eta_max = eta (CDomain.cpp:294-397 semantics)."""
    bed = np.asarray(bed, dtype=np.float64)
    st = np.zeros(bed.shape + (4,), dtype=dtype)           # built in the target precision: no second full-size copy
    st[..., 0] = bed + depth                               # (sum in double, rounded once on assignment)
    st[..., 1] = st[..., 0]
    if qx is not None:
        st[..., 2] = qx
    if qy is not None:
        st[..., 3] = qy
    return st


def fractal_dem(rows, cols, seed, amplitude=50.0, hurst=0.8, base=64):
    """Cheap multi-octave terrain: bilinearly upsampled random lattices, persistence 2^-H.

    O(rows*cols) memory and time, deterministic for a given seed, works at 32768^2.
    """
    rng = np.random.default_rng(seed)
    z = np.zeros((rows, cols), dtype=np.float64)
    amp, cell, total = 1.0, max(rows, cols) / 2.0, 0.0
    while cell >= max(2.0, min(rows, cols) / 4096.0 * 2.0) and cell >= base / 16.0:
        ny, nx = int(np.ceil(rows / cell)) + 2, int(np.ceil(cols / cell)) + 2
        lat = rng.uniform(-1.0, 1.0, size=(ny, nx))
        yy = (np.arange(rows) + 0.5) / cell
        xx = (np.arange(cols) + 0.5) / cell
        y0, x0 = np.floor(yy).astype(np.int64), np.floor(xx).astype(np.int64)
        fy, fx = (yy - y0)[:, None], (xx - x0)[None, :]
        fy, fx = fy * fy * (3 - 2 * fy), fx * fx * (3 - 2 * fx)
        top = lat[y0][:, x0] * (1 - fx) + lat[y0][:, x0 + 1] * fx
        bot = lat[y0 + 1][:, x0] * (1 - fx) + lat[y0 + 1][:, x0 + 1] * fx
        z += amp * (top * (1 - fy) + bot * fy)
        total += amp
        amp *= 2.0 ** (-hurst)
        cell /= 2.0
    z = (z / total) * amplitude + amplitude
    return round4(z)


def dam_break(n, inner_level=10.0, outer_level=1.0, radius=None, dtype=np.float64):
    """Config C2: circular dam break on a flat bed, n x n cells."""
    radius = n / 8.0 if radius is None else radius
    y, x = np.mgrid[0:n, 0:n]
    r2 = (x - n / 2.0 + 0.5) ** 2 + (y - n / 2.0 + 0.5) ** 2
    bed = np.zeros((n, n))
    depth = np.where(r2 < radius * radius, inner_level, outer_level)
    return bed.astype(dtype), make_states(bed, depth, dtype=dtype), np.full((n, n), 0.03, dtype=dtype)


def pluvial(rows, cols, seed=20260817, manning=0.035, dtype=np.float64, wet_fraction=0.0):
    """Configs C3/C4: dry fractal terrain (optionally pre-wetted hollows for early flow)."""
    bed = fractal_dem(rows, cols, seed)
    depth = np.zeros_like(bed)
    if wet_fraction > 0.0:
        level = np.quantile(bed[:: max(1, rows // 512), :: max(1, cols // 512)], wet_fraction)
        depth = round4(np.maximum(level - bed, 0.0))
    return bed.astype(dtype), make_states(bed, depth, dtype=dtype), np.full(bed.shape, manning, dtype=dtype)


def river_valley(rows, cols, seed=20260819, manning=0.03, dtype=np.float64):
    """Config C5: valley falling west->east with a 2 m deep river along the centre line."""
    y, x = np.mgrid[0:rows, 0:cols].astype(np.float64)
    mid, width = rows / 2.0, max(8.0, rows / 16.0)
    rng = np.random.default_rng(seed)
    noise = 0.5 * rng.uniform(-1.0, 1.0, size=(rows, cols))
    bed = round4(0.001 * (cols - x) + 20.0 * (1.0 - np.exp(-(((y - mid) / width) ** 2))) + noise + 1.0)
    thalweg = 0.001 * (cols - x) + 1.0
    depth = round4(np.maximum(thalweg + 2.0 - bed, 0.0))
    return bed.astype(dtype), make_states(bed, depth, dtype=dtype), np.full(bed.shape, manning, dtype=dtype)


def lake_at_rest(rows, cols, seed=3, level=None, dtype=np.float64):
    """Known-answer case (tools/model-builder/tests/TestLakeAtRest.js:59-70): still water over
    uneven, partly emerged terrain must stay still."""
    bed = fractal_dem(rows, cols, seed, amplitude=5.0)
    level = float(np.median(bed)) if level is None else level
    depth = np.maximum(level - bed, 0.0)
    return bed.astype(dtype), make_states(bed, depth, dtype=dtype), np.full(bed.shape, 0.03, dtype=dtype)


def random_wet_dry(rows, cols, seed, dtype=np.float64, disabled=True):
    """Adversarial parity input: rough bed, patchy water, random discharges, a few disabled
    cells (eta_max = -9999, src/Schemes/CLSchemeGodunov.clc:214)."""
    rng = np.random.default_rng(seed)
    bed = round4(rng.uniform(0.0, 2.0, size=(rows, cols)) + fractal_dem(rows, cols, seed + 1, amplitude=1.5))
    depth = np.where(rng.uniform(size=bed.shape) < 0.6, rng.uniform(0.0, 1.5, size=bed.shape), 0.0)
    depth = round4(np.where(rng.uniform(size=bed.shape) < 0.05, 1e-6 * rng.uniform(size=bed.shape), depth))
    wet = depth > 0
    qx = np.where(wet, rng.normal(0.0, 0.4, size=bed.shape), 0.0)
    qy = np.where(wet, rng.normal(0.0, 0.4, size=bed.shape), 0.0)
    st = make_states(bed, depth, qx, qy, dtype=np.float64)
    if disabled:
        off = rng.uniform(size=bed.shape) < 0.01
        st[off, 0] = -9999.0
        st[off, 1] = -9999.0
        st[off, 2:] = 0.0
    man = rng.uniform(0.01, 0.06, size=bed.shape)
    return bed.astype(dtype), st.astype(dtype), man.astype(dtype)
