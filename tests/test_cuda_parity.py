"""Parity of the CUDA executor (through the C ABI) with the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north star): max |d eta| <= 1e-9 m in fp64, <= 1e-4 m in fp32, after a
fixed number of iterations, with identical wet-cell counts and identical timestep counts.
With HP_OPT_STRICT_FP (no FMA contraction) friction-free Godunov / MUSCL-Hancock runs are
bit-identical to the oracle; with friction the only difference is pow().
"""
import numpy as np
import pytest

from hipims_ocl_b200 import config as hc
from hipims_ocl_b200 import executor as hx
from oracle import cpu_sim
from tests.helpers import add_standard_boundaries, dtype_of, make_cfg, scenario

pytestmark = pytest.mark.gpu

TOL = {"double": 1e-9, "single": 1e-4}


@pytest.fixture(scope="module")
def ex():
    e = hx.Executor(0)
    yield e
    e.close()


def wet_count(states, bed, eps=1e-10):
    return int(((states[..., 0] - bed) > eps).sum())


def run_pair(ex, cfg, scen, bdy, iters, options, chunks=None):
    dt = dtype_of(cfg.precision)
    bed, st, man = scenario(scen, cfg.rows, cfg.cols, dt)
    orc = cpu_sim.CpuSim("oracle", cfg)
    gpu = hx.CudaScheme(ex, cfg, options=options)
    for sim in (orc, gpu):
        sim.upload(st, bed, man)
        add_standard_boundaries(sim, cfg, bdy)
        sim.set_target(1.0e6)
    for n in (chunks or [iters]):
        orc.iterate(n)
        gpu.iterate(n)
    return orc, gpu, bed


def compare(orc, gpu, bed, cfg, exact=False):
    so, sg = orc.stats(), gpu.stats()
    assert sg["batch_successful"] == so["batch_successful"]
    assert sg["batch_skipped"] == so["batch_skipped"]
    assert sg["use_alternate"] == so["use_alternate"]
    a_o, b_o = orc.download_both()
    a_g, b_g = gpu.download_both()
    if exact:
        assert sg == so
        if cfg.scheme == hc.SCHEME_MUSCL_HANCOCK:
            # the reference updates MUSCL-Hancock in place; the fused kernel ping-pongs, so only
            # the current buffer is comparable
            np.testing.assert_array_equal(gpu.download(), orc.download())
        else:
            np.testing.assert_array_equal(a_g, a_o)
            np.testing.assert_array_equal(b_g, b_o)
        return
    tol = TOL[cfg.precision]
    rel = 1e-9 if cfg.precision == "double" else 1e-4
    assert abs(sg["time"] - so["time"]) <= rel * max(1.0, abs(so["time"]))
    assert abs(sg["timestep"] - so["timestep"]) <= rel * max(1.0, abs(so["timestep"]))
    cur_o, cur_g = orc.download(), gpu.download()
    assert np.isfinite(cur_g).all()
    assert np.abs(cur_g[..., 0] - cur_o[..., 0]).max() <= tol            # eta
    assert np.abs(cur_g[..., 1] - cur_o[..., 1]).max() <= tol            # eta_max
    assert np.abs(cur_g[..., 2:] - cur_o[..., 2:]).max() <= 100 * tol    # discharges
    assert wet_count(cur_g, bed) == wet_count(cur_o, bed)
    vol_o = (cur_o[..., 0].astype(np.float64) - bed).sum()
    vol_g = (cur_g[..., 0].astype(np.float64) - bed).sum()
    assert abs(vol_g - vol_o) <= (1e-10 if cfg.precision == "double" else 1e-5) * max(1.0, abs(vol_o))
    if cfg.scheme != hc.SCHEME_MUSCL_HANCOCK:  # the other ping-pong buffer too (stale-dst rule, Q2)
        other_o, other_g = (a_o, a_g) if so["use_alternate"] else (b_o, b_g)
        assert np.abs(other_g[..., 0] - other_o[..., 0]).max() <= tol


STRICT_EXACT = [
    ("godunov", "double", "dambreak", "none", 64, 60, {"friction": False}),
    ("godunov", "double", "dambreak-dry", "none", 64, 60, {"friction": False}),
    ("godunov", "single", "dambreak", "none", 64, 60, {"friction": False}),
    ("godunov", "double", "wetdry", "rain", 45, 80, {"friction": False}),
    ("godunov", "double", "wetdry", "rain", 45, 80, {"friction": False, "quirks": 0}),
    ("godunov", "double", "pluvial-wet", "gridded", 48, 150, {"friction": False}),
    ("godunov", "double", "dambreak", "none", 48, 40, {"friction": False, "dynamic": False, "fixed_dt": 0.01}),
    ("muscl-hancock", "double", "dambreak", "none", 64, 50, {"friction": False}),
    ("muscl-hancock", "double", "dambreak-dry", "none", 64, 50, {"friction": False}),
    ("muscl-hancock", "single", "wetdry", "rain", 50, 60, {"friction": False}),
]


@pytest.mark.parametrize("scheme,precision,scen,bdy,n,iters,extra", STRICT_EXACT)
def test_strict_mode_is_bit_identical(ex, scheme, precision, scen, bdy, n, iters, extra):
    cfg = make_cfg(scheme, precision, n, n, **extra)
    orc, gpu, bed = run_pair(ex, cfg, scen, bdy, iters, hx.OPT_STRICT_FP, chunks=[1, 2, 5, iters - 8])
    compare(orc, gpu, bed, cfg, exact=True)


TOLERANCE_CASES = [
    ("godunov", "double", "dambreak", "none", 96, 200, {}),
    ("godunov", "double", "dambreak-dry", "none", 96, 200, {}),
    ("godunov", "single", "dambreak", "none", 96, 200, {}),
    # thin-film rain cases: the fixed step count is shorter because every implementation that does
    # not share the reference's pow() bit for bit drifts apart through the scheme's discontinuous
    # thresholds (|D| < eps => 0, h < 1e-5 => first order); see DESIGN.md "Parity"
    ("godunov", "double", "pluvial", "rain+loss", 50, 80, {"delta": 2.0}),
    ("godunov", "double", "valley", "cells", 64, 200, {}),
    ("godunov", "double", "pluvial-wet", "gridded", 64, 200, {}),
    ("godunov", "double", "lake", "none", 64, 100, {}),
    ("inertial", "double", "pluvial-wet", "rain", 64, 200, {}),
    ("inertial", "single", "valley", "cells", 64, 200, {}),
    ("muscl-hancock", "double", "dambreak", "none", 96, 200, {}),
    ("muscl-hancock", "double", "dambreak-dry", "none", 96, 200, {}),
    ("muscl-hancock", "single", "dambreak", "none", 96, 200, {}),
    ("muscl-hancock", "double", "valley", "cells", 64, 200, {}),
    ("muscl-hancock", "double", "pluvial-wet", "rain", 64, 40, {}),
    # wide enough for whole 32-column strips to be exactly dry (the marching kernels copy such rows through) while the
    # dam-break front runs into them
    ("muscl-hancock", "double", "dambreak-dry", "none", 160, 150, {}),
    ("muscl-hancock", "single", "dambreak-dry", "none", 160, 150, {}),
    ("godunov", "double", "dambreak-dry", "none", 160, 150, {}),
    ("inertial", "double", "dambreak-dry", "none", 160, 150, {}),
]


# "fast" runs the default kernels (Godunov: TMA tiles, MUSCL-Hancock: marching warps, inertial: wide marching warps
# with two columns per lane), "fast-other" the other
# kernel of each scheme (Godunov marching, MUSCL-Hancock tiles, inertial one column per lane); "fast-narrow" /
# "fast-wide" pin the marching kernels to one / two columns per lane whatever the default of the precision is
OTHER_KERNELS = hx.OPT_TILE_KERNELS | hx.OPT_MARCH_GODUNOV | hx.OPT_NARROW_MARCH
WIDTHS = [hx.OPT_NARROW_MARCH, hx.OPT_WIDE_MARCH | hx.OPT_MARCH_GODUNOV]


@pytest.mark.parametrize("options", [hx.OPT_STRICT_FP, 0, hx.OPT_NO_GRAPH, OTHER_KERNELS] + WIDTHS,
                         ids=["strict", "fast", "fast-nograph", "fast-other", "fast-narrow", "fast-wide"])
@pytest.mark.parametrize("scheme,precision,scen,bdy,n,iters,extra", TOLERANCE_CASES)
def test_parity_within_tolerance(ex, options, scheme, precision, scen, bdy, n, iters, extra):
    cfg = make_cfg(scheme, precision, n, n, **extra)
    orc, gpu, bed = run_pair(ex, cfg, scen, bdy, iters, options)
    compare(orc, gpu, bed, cfg)


SINGLE_STEP = [(s, p) for s in ("godunov", "muscl-hancock", "inertial") for p in ("double", "single")]


@pytest.mark.parametrize("options", [hx.OPT_STRICT_FP, 0, OTHER_KERNELS] + WIDTHS, ids=["strict", "fast", "fast-other", "fast-narrow", "fast-wide"])
@pytest.mark.parametrize("scheme,precision", SINGLE_STEP)
def test_single_iteration_on_adversarial_input(ex, options, scheme, precision):
    """One iteration from identical adversarial states (rough bed, wet/dry patches, disabled cells,
    friction on): rounding cannot be amplified yet, so every flavour must agree to a few ulp."""
    rows, cols = 70, 93
    cfg = make_cfg(scheme, precision, rows, cols)
    tol = 2e-12 if precision == "double" else 2e-5
    for seed in (21, 22, 23):
        bed, st, man = scenario("wetdry", rows, cols, dtype_of(precision), seed=seed)
        orc = cpu_sim.CpuSim("oracle", cfg)
        gpu = hx.CudaScheme(ex, cfg, options=options)
        for sim in (orc, gpu):
            sim.upload(st, bed, man)
            sim.set_target(1e6)
            sim.set_clock(5.0, 0.02, 0.0)
            sim.iterate(1)
        a_o, b_o = orc.download_both()
        a_g, b_g = gpu.download_both()
        cur_o, cur_g = orc.download(), gpu.download()
        scale = np.maximum(1.0, np.abs(cur_o))
        assert (np.abs(cur_g - cur_o) / scale).max() <= tol
        if scheme != "muscl-hancock":
            assert (np.abs(a_g - a_o) / np.maximum(1.0, np.abs(a_o))).max() <= tol
            assert (np.abs(b_g - b_o) / np.maximum(1.0, np.abs(b_o))).max() <= tol
        so, sg = orc.stats(), gpu.stats()
        assert abs(sg["timestep"] - so["timestep"]) <= tol * max(1.0, abs(so["timestep"]))
        gpu.close()


@pytest.mark.parametrize("scheme", ["godunov", "muscl-hancock"])
def test_long_thin_film_run_stays_close(ex, scheme):
    """400 iterations of rain on dry terrain (the Newcastle-like config): drift stays bounded,
    volume and wet-cell counts agree."""
    n = 50
    cfg = make_cfg(scheme, "double", n, n, delta=2.0)
    orc, gpu, bed = run_pair(ex, cfg, "pluvial", "rain+loss", 400, 0)
    so, sg = orc.stats(), gpu.stats()
    assert sg["batch_successful"] == so["batch_successful"] == 400
    cur_o, cur_g = orc.download(), gpu.download()
    assert np.abs(cur_g[..., 0] - cur_o[..., 0]).max() <= 1e-5
    vol_o, vol_g = (cur_o[..., 0] - bed).sum(), (cur_g[..., 0] - bed).sum()
    assert abs(vol_g - vol_o) <= 1e-5 * vol_o
    assert abs(wet_count(cur_g, bed) - wet_count(cur_o, bed)) <= 2


def test_ragged_sizes_and_ring(ex):
    """Sizes that are not multiples of the tile, and the frozen outer ring."""
    for rows, cols in ((3, 3), (5, 67), (37, 4), (33, 65)):
        cfg = make_cfg("godunov", "double", rows, cols, friction=False)
        bed, st, man = scenario("wetdry", rows, cols, np.float64, seed=rows * 100 + cols)
        orc = cpu_sim.CpuSim("oracle", cfg)
        gpu = hx.CudaScheme(ex, cfg, options=hx.OPT_STRICT_FP)
        for sim in (orc, gpu):
            sim.upload(st, bed, man)
            sim.set_target(1e6)
            sim.iterate(7)
        np.testing.assert_array_equal(gpu.download(), orc.download())
        out = gpu.download()
        np.testing.assert_array_equal(out[0], st[0])
        np.testing.assert_array_equal(out[:, 0], st[:, 0])


@pytest.mark.parametrize("options", [0, OTHER_KERNELS] + WIDTHS, ids=["fast", "fast-other", "fast-narrow", "fast-wide"])
@pytest.mark.parametrize("scheme", ["godunov", "muscl-hancock", "inertial"])
def test_ragged_sizes_fast_kernels(ex, scheme, options):
    """Sizes around the tile (32 x 8) and strip (28 / 30 / 60 columns, 4 warps per CTA) boundaries of the fast kernels, a
    few iterations from adversarial states: every flavour must agree with the oracle to rounding, ring frozen."""
    for rows, cols in ((3, 3), (4, 31), (5, 67), (37, 4), (33, 65), (9, 30), (12, 61), (7, 121), (66, 29), (6, 60), (11, 241), (5, 59)):
        cfg = make_cfg(scheme, "double", rows, cols)
        bed, st, man = scenario("wetdry", rows, cols, np.float64, seed=rows * 100 + cols)
        orc = cpu_sim.CpuSim("oracle", cfg)
        gpu = hx.CudaScheme(ex, cfg, options=options)
        for sim in (orc, gpu):
            sim.upload(st, bed, man)
            sim.set_target(1e6)
            sim.set_clock(5.0, 0.02, 0.0)
            sim.iterate(3)
        want, got = orc.download(), gpu.download()
        # (the 3 x 3 terrain generator degenerates to NaN beds: both sides must then agree on where the NaNs are)
        np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-11, equal_nan=True, err_msg=str((rows, cols)))
        np.testing.assert_array_equal(got[0], st[0])
        np.testing.assert_array_equal(got[:, 0], st[:, 0])
        np.testing.assert_array_equal(got[-1], st[-1])
        np.testing.assert_array_equal(got[:, -1], st[:, -1])
        dt_o, dt_g = orc.stats()["timestep"], gpu.stats()["timestep"]
        assert (np.isnan(dt_o) and np.isnan(dt_g)) or abs(dt_g - dt_o) <= 1e-11 * max(1.0, abs(dt_o))
        gpu.close()


def test_sync_point_suspension_and_update(ex):
    n = 48
    cfg = make_cfg("godunov", "double", n, n, friction=False)
    bed, st, man = scenario("dambreak", n, n, np.float64)
    orc = cpu_sim.CpuSim("oracle", cfg)
    gpu = hx.CudaScheme(ex, cfg, options=hx.OPT_STRICT_FP)
    for sim in (orc, gpu):
        sim.upload(st, bed, man)
        sim.set_target(0.25)
        sim.iterate(30)
    assert gpu.stats() == orc.stats()
    assert gpu.stats()["timestep"] < 0 and gpu.stats()["batch_skipped"] > 0
    for sim in (orc, gpu):
        sim.set_target(0.5)
        sim.update_timestep()
        sim.reset_counters()
        sim.iterate(10)
    assert gpu.stats() == orc.stats()
    np.testing.assert_array_equal(gpu.download(), orc.download())


@pytest.mark.parametrize("options", [hx.OPT_STRICT_FP, 0, hx.OPT_MARCH_GODUNOV | hx.OPT_NARROW_MARCH, hx.OPT_MARCH_GODUNOV | hx.OPT_WIDE_MARCH],
                         ids=["strict", "tiles", "march", "wide"])
def test_godunov_dt0_keep_rule(ex, options):
    """HP_QUIRK_GODUNOV_DT0_KEEP: once the clock sits at its target (timestep <= 0) a Godunov step leaves the destination
    untouched -- gts_cacheEnabled's rule (CLSchemeGodunov.clc:477-478) -- instead of copying the source through like the
    default gts_cacheDisabled (:201-206); both ping-pong buffers then stay what they were, the older one stale."""
    n = 96
    cfg = make_cfg("godunov", "double", n, n, friction=False)
    cfg.quirks |= hc.QUIRK_GODUNOV_DT0_KEEP
    bed, st, man = scenario("dambreak", n, n, np.float64)
    orc = cpu_sim.CpuSim("oracle", cfg)
    gpu = hx.CudaScheme(ex, cfg, options=options)
    for sim in (orc, gpu):
        sim.upload(st, bed, man)
        sim.set_target(0.25)
        sim.iterate(33)                     # reaches the target after ~20 iterations; an odd count on purpose
    so, sg = orc.stats(), gpu.stats()
    assert sg["timestep"] < 0 and sg["batch_skipped"] == so["batch_skipped"] > 2
    (ao, bo), (ag, bg) = orc.download_both(), gpu.download_both()
    assert not np.array_equal(ao, bo)                                   # the older buffer was left alone
    if options & hx.OPT_STRICT_FP:
        np.testing.assert_array_equal(ag, ao)
        np.testing.assert_array_equal(bg, bo)
    else:
        assert np.abs(ag - ao).max() <= 1e-9 and np.abs(bg - bo).max() <= 1e-9
    for sim in (orc, gpu):                                              # moving the target on resumes from the right buffer
        sim.set_target(0.5)
        sim.update_timestep()
        sim.reset_counters()
        sim.iterate(10)
    so, sg = orc.stats(), gpu.stats()
    assert (sg["batch_successful"], sg["batch_skipped"]) == (so["batch_successful"], so["batch_skipped"]) and so["batch_successful"] >= 5
    assert np.abs(gpu.download() - orc.download()).max() <= (0 if options & hx.OPT_STRICT_FP else 1e-9)
    gpu.close(); orc.close()


def test_link_rows_roundtrip(ex):
    """hp_scheme_read_rows / write_rows (CDomainLink pull/push)."""
    cfg = make_cfg("godunov", "double", 20, 31)
    bed, st, man = scenario("wetdry", 20, 31, np.float64, seed=3)
    gpu = hx.CudaScheme(ex, cfg)
    gpu.upload(st, bed, man)
    np.testing.assert_array_equal(gpu.read_rows(5, 4), st[5:9])
    patch = st[5:9] + 1.0
    gpu.write_rows(5, patch)
    out = gpu.download()
    np.testing.assert_array_equal(out[5:9], patch)
    np.testing.assert_array_equal(out[:5], st[:5])


def test_size_independent_properties_at_scale(ex):
    """At a size the oracle would take minutes for: lake at rest stays at rest, and the dam break
    conserves volume and keeps its 4-fold symmetry."""
    n = 2048
    cfg = make_cfg("godunov", "double", n, n)
    bed, st, man = scenario("lake", n, n, np.float64)
    gpu = hx.CudaScheme(ex, cfg)
    gpu.upload(st, bed, man)
    gpu.set_target(1e6)
    gpu.iterate(50)
    out = gpu.download()
    wet = (st[..., 0] - bed) > 1e-10
    assert np.abs(out[..., 0][wet] - st[..., 0][wet]).max() < 1e-9
    assert np.abs(out[..., 2:]).max() < 1e-7
    gpu.close()

    bed, st, man = scenario("dambreak", n, n, np.float64)
    gpu = hx.CudaScheme(ex, cfg.with_(friction=False))
    gpu.upload(st, bed, man)
    gpu.set_target(1e6)
    gpu.iterate(100)
    out = gpu.download()
    v0, v1 = (st[..., 0] - bed).sum(), (out[..., 0] - bed).sum()
    assert abs(v1 - v0) / v0 < 1e-12
    eta = out[..., 0]
    assert np.abs(eta - eta[::-1, :]).max() < 1e-9 and np.abs(eta - eta[:, ::-1]).max() < 1e-9
    assert np.abs(eta - eta.T).max() < 1e-9
    assert gpu.stats()["batch_successful"] == 100


def test_handles_are_thread_safe(ex):
    """The reference's main thread polls readKeyStatistics and reads cells back while the scheme's worker thread enqueues
    batches (src/Schemes/CSchemeGodunov.cpp:1116-1141, CScheme.h:137-139): every entry point takes the handle's lock, so
    concurrent use is safe and the result is the one of the same iterations issued serially."""
    import threading
    n = 256
    cfg = make_cfg("muscl-hancock", "double", n, n)
    bed, st, man = scenario("dambreak", n, n, np.float64)
    serial = hx.CudaScheme(ex, cfg)
    serial.upload(st, bed, man)
    serial.set_target(1.0e6)
    serial.iterate(240)
    want, want_stats = serial.download(), serial.stats()
    serial.close()

    sim = hx.CudaScheme(ex, cfg)
    sim.upload(st, bed, man)
    sim.set_target(1.0e6)
    errors, seen = [], []

    def worker():
        try:
            for _ in range(60):
                sim.iterate(4, sync=False)
        except Exception as e:            # pragma: no cover
            errors.append(e)

    t = threading.Thread(target=worker)
    t.start()
    while t.is_alive():
        s = sim.stats()                   # synchronous read of the device clock, racing the enqueues
        seen.append((s["batch_successful"], s["time"]))
        rows = sim.read_rows(n // 2, 2)   # CDomainLink-style partial read-back in the middle of a batch
        assert np.isfinite(rows).all()
    t.join()
    sim.sync()
    assert not errors, errors
    assert all(b[0] >= a[0] and b[1] >= a[1] for a, b in zip(seen, seen[1:]))     # the clock only moves forward
    assert sim.stats() == want_stats
    np.testing.assert_array_equal(sim.download(), want)
    sim.close()
