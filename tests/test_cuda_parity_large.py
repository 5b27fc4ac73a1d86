"""Oracle parity of the DEFAULT fast kernels at sizes where their persistent loops wrap.

The marching kernels (MUSCL-Hancock, inertial, Godunov-march) deal `march_runs` interleaved runs of rows to every
CTA once a CTA has >= 128 (strip group, row) units (hp_march_kernels.cuh: march_runs): ring slots, mbarrier phases
and the carried row state then cross run boundaries.  The Godunov tile kernel re-arms its mbarrier and re-issues its
TMA boxes once there are more tiles than resident CTAs (6 x 148 in fp64, 9 x 148 in fp32).  The small parity cases
(<= 160 x 160) never reach either path; every number bench.py quotes comes from them.  Here the bench's own
workload generators (BASELINE configs[1], [2], [4], cropped) run 12 iterations against the CPU oracle
(src/Schemes/CLSchemeMUSCLHancock.clc:533-801, CLSchemeInertial.clc:27-163, CLSchemeGodunov.clc:164-384 restated),
with the north-star tolerances, identical wet-cell and timestep counts.
"""
import numpy as np
import pytest

import bench
from hipims_ocl_b200 import config as hc
from hipims_ocl_b200 import executor as hx
from oracle import cpu_sim

pytestmark = pytest.mark.gpu

TOL = {"double": 1e-9, "single": 1e-4}


@pytest.fixture(scope="module")
def ex():
    e = hx.Executor(0)
    yield e
    e.close()


def crop(workload, rows, cols):
    w = dict(bench.WORKLOADS[workload])
    w.update(cols=cols, rows_per_gpu=rows)
    return w


def run_both(ex, w, rows, cols, iters, options=0, warm_time=None, hydro=0.97):
    cfg = bench.cfg_for(w, rows, cols)
    dtype = np.float64 if cfg.precision == "double" else np.float32
    bed, st, man = bench.make_inputs(w, rows, cols, dtype)
    orc = cpu_sim.CpuSim("oracle", cfg)
    gpu = hx.CudaScheme(ex, cfg, options=options)
    for sim in (orc, gpu):
        sim.upload(st, bed, man)
        bench.attach_boundaries(sim, w, cols, rows)
        sim.set_target(1.0e7)
        if warm_time is not None:               # start inside the forcing series (rain falling, river flowing), with the
            sim.set_clock(warm_time, cfg.initial_dt, hydro)  # hydrological accumulator about to fire
        sim.iterate(iters)
    return cfg, orc, gpu, bed, st


def check(cfg, orc, gpu, bed, st, iters):
    so, sg = orc.stats(), gpu.stats()
    assert sg["batch_successful"] == so["batch_successful"] == iters
    assert sg["batch_skipped"] == so["batch_skipped"]
    rel = 1e-9 if cfg.precision == "double" else 1e-4
    assert abs(sg["time"] - so["time"]) <= rel * max(1.0, abs(so["time"]))
    assert abs(sg["timestep"] - so["timestep"]) <= rel * max(1.0, abs(so["timestep"]))
    cur_o, cur_g = orc.download(), gpu.download()
    tol = TOL[cfg.precision]
    assert np.isfinite(cur_g).all()
    assert not np.array_equal(cur_o[..., 0], st[..., 0])          # the run did something
    assert np.abs(cur_g[..., 0] - cur_o[..., 0]).max() <= tol
    assert np.abs(cur_g[..., 1] - cur_o[..., 1]).max() <= tol
    assert np.abs(cur_g[..., 2:] - cur_o[..., 2:]).max() <= 100 * tol
    wet_o = int(((cur_o[..., 0] - bed) > 1e-10).sum())
    wet_g = int(((cur_g[..., 0] - bed) > 1e-10).sum())
    assert wet_g == wet_o
    vol_o = (cur_o[..., 0].astype(np.float64) - bed).sum()
    vol_g = (cur_g[..., 0].astype(np.float64) - bed).sum()
    assert abs(vol_g - vol_o) <= (1e-10 if cfg.precision == "double" else 1e-5) * max(1.0, abs(vol_o))
    if cfg.scheme != hc.SCHEME_MUSCL_HANCOCK:                      # stale-destination rule
        a_o, b_o = orc.download_both()
        a_g, b_g = gpu.download_both()
        oth_o, oth_g = (a_o, a_g) if so["use_alternate"] else (b_o, b_g)
        assert np.abs(oth_g[..., 0] - oth_o[..., 0]).max() <= tol


MARCH = hx.OPT_MARCH_GODUNOV
NARROW = hx.OPT_NARROW_MARCH                 # one column per lane where the default is the two-column ("wide") kernel
WIDE = hx.OPT_WIDE_MARCH                     # two columns per lane where the default is the one-column kernel

# (workload, rows, cols, options, start time): rows x cols chosen so that march_runs >= 2 for that kernel's grid
WRAP_CASES = [
    pytest.param("dambreak4096-mh", 3072, 4096, 0, None, id="mh-f64-dambreak"),
    pytest.param("dambreak4096-mh-f32", 3072, 4096, 0, None, id="mh-f32-dambreak"),
    pytest.param("river32768", 4096, 4096, 0, 700.0, id="mh-f64-river-cells"),                # configs[4] cropped
    pytest.param("dambreak4096-mh", 3072, 4096, WIDE, None, id="mh-wide-f64-dambreak"),
    pytest.param("river32768", 4096, 4096, WIDE, 700.0, id="mh-wide-f64-river-cells"),
    pytest.param("dambreak4096-mh", 3072, 4096, NARROW, None, id="mh-narrow-f64-dambreak"),
    pytest.param("dambreak4096-mh-f32", 3072, 4096, NARROW, None, id="mh-narrow-f32-dambreak"),
    pytest.param("river32768", 4096, 4096, NARROW, 700.0, id="mh-narrow-f64-river-cells"),
    pytest.param("dambreak4096-inertial", 4096, 4096, 0, None, id="inertial-f64-dambreak"),
    pytest.param("dambreak4096-inertial-f32", 4096, 4096, 0, None, id="inertial-f32-dambreak"),
    pytest.param("dambreak4096-inertial", 4096, 4096, NARROW, None, id="inertial-narrow-f64-dambreak"),
    pytest.param("dambreak4096-inertial-f32", 4096, 4096, NARROW, None, id="inertial-narrow-f32-dambreak"),
    pytest.param("dambreak4096", 3072, 4096, MARCH | NARROW, None, id="godunov-march-f64-dambreak"),
    pytest.param("dambreak4096-f32", 3072, 4096, MARCH | NARROW, None, id="godunov-march-f32-dambreak"),
    pytest.param("dambreak4096", 4096, 4096, MARCH | WIDE, None, id="godunov-wide-f64-dambreak"),
    pytest.param("dambreak4096-f32", 4096, 4096, MARCH | WIDE, None, id="godunov-wide-f32-dambreak"),
    # the tile kernel: 1024 x 1536 = 32 x 192 = 6144 tiles > 888 (fp64) / 1332 (fp32) resident CTAs
    pytest.param("dambreak4096", 1024, 1536, 0, None, id="godunov-tiles-f64-dambreak"),
    pytest.param("dambreak4096-f32", 1024, 1536, 0, None, id="godunov-tiles-f32-dambreak"),
]


@pytest.mark.parametrize("workload,rows,cols,options,t0", WRAP_CASES)
def test_wrapping_kernels_match_the_oracle(ex, workload, rows, cols, options, t0):
    iters = 12
    w = crop(workload, rows, cols)
    cfg, orc, gpu, bed, st = run_both(ex, w, rows, cols, iters, options, warm_time=t0)
    check(cfg, orc, gpu, bed, st, iters)
    gpu.close()
    orc.close()


@pytest.mark.parametrize("scheme,options", [("godunov", 0), ("godunov", MARCH | NARROW), ("godunov", MARCH | WIDE), ("muscl-hancock", WIDE),
                                            ("muscl-hancock", NARROW), ("inertial", 0), ("inertial", NARROW)],
                         ids=["godunov-tiles", "godunov-march", "godunov-wide", "mh-wide", "mh-narrow", "inertial-wide", "inertial-narrow"])
def test_wrapping_kernels_on_wet_dry_terrain(ex, scheme, options):
    """Random rough terrain with wet and dry patches, fronts everywhere (the adversarial generator of the small parity
    cases) at a size where the persistent loops wrap: every dry-side / stop-flag / stale-destination branch next to a
    tile or run boundary."""
    from tests.helpers import make_cfg, scenario
    rows, cols, iters = 4608, 3072, 3
    cfg = make_cfg(scheme, "double", rows, cols)
    bed, st, man = scenario("wetdry", rows, cols, np.float64, seed=77)
    orc = cpu_sim.CpuSim("oracle", cfg)
    gpu = hx.CudaScheme(ex, cfg, options=options)
    for sim in (orc, gpu):
        sim.upload(st, bed, man)
        sim.set_target(1.0e7)
        sim.iterate(iters)
    check(cfg, orc, gpu, bed, st, iters)
    gpu.close()
    orc.close()


def test_sheet_flow_deviations_are_isolated(ex):
    """configs[2] cropped, the first iteration applying 40 s worth of rain at once (hydrological accumulator preset): a
    0.6 mm sheet running down every slope, friction dominated, every cell next to one of the scheme's discontinuous
    switches (reconstructed depth <= eps => no velocity, stop flags, |D| < eps => 0).  A rounding difference that
    flips one switches the cell's momentum by ~1e-5 m2/s, so agreement to 1e-9 m is a statement about ALL cells only
    for a bit-identical implementation: the strict flavour (whose only difference from the oracle is the last bit of
    pow()) has 4 cells of 1 Mi beyond 1e-9 m after 12 iterations, 5e-8 m at worst; the fast kernels, with FMA
    contraction and reciprocal-based division in every term, 64 cells and 4e-7 m (tools/diag_large.py).  The bar here:
    such cells stay isolated (< 2e-4 of the domain) and small (a hundredth of the film), every global figure agrees."""
    rows, cols, iters = 3072, 4096, 12
    w = crop("pluvial16384", rows, cols)
    cfg, orc, gpu, bed, st = run_both(ex, w, rows, cols, iters, 0, warm_time=100.0, hydro=40.0)
    want, got = orc.download(), gpu.download()
    so, sg = orc.stats(), gpu.stats()
    gpu.close(); orc.close()
    dev = np.abs(got[..., 0] - want[..., 0])
    assert np.quantile((want[..., 0] - st[..., 0])[2:-2, 2:-2], 0.01) > 4.0e-4   # the sheet is (nearly) everywhere
    assert (dev > 1e-9).mean() < 2e-4 and dev.max() < 5e-6
    assert np.median(dev) <= 1e-13
    assert sg["batch_successful"] == so["batch_successful"] == iters
    assert abs(sg["timestep"] - so["timestep"]) <= 1e-9 * so["timestep"]
    assert int(((got[..., 0] - bed) > 1e-10).sum()) == int(((want[..., 0] - bed) > 1e-10).sum())
    vol_o, vol_g = (want[..., 0] - bed).sum(), (got[..., 0] - bed).sum()
    assert abs(vol_g - vol_o) <= 1e-10 * vol_o


def test_rain_film_deviation_is_that_of_a_bit_faithful_run(ex):
    """configs[2] cropped, with the rain firing in the first iteration (hydrological accumulator at 0.97 s): every dry
    cell gets a 1.4e-5 m film, right above the scheme's `h < 1e-5 => first order` switch.  From there on NO implementation
    that does not share the reference's pow() bit for bit can stay within 1e-9 m: the strict flavour (reference operation
    order, IEEE division and roots, bit-identical to the oracle wherever pow() is not involved) is 7e-9 away after two
    more iterations and 3e-8 after twelve (tools/diag_large.py).  So the bar for the fast kernels here is the deviation
    of that bit-faithful witness: same order of magnitude, identical timestep, wet-cell count and volume."""
    rows, cols, iters = 3072, 4096, 12
    w = crop("pluvial16384", rows, cols)
    cfg, orc, fast, bed, st = run_both(ex, w, rows, cols, iters, 0, warm_time=100.0)
    want = orc.download()
    got = fast.download()
    so, sg = orc.stats(), fast.stats()
    fast.close()
    cfg2, orc2, strict, _, _ = run_both(ex, w, rows, cols, iters, hx.OPT_STRICT_FP, warm_time=100.0)
    witness = strict.download()
    strict.close(); orc2.close(); orc.close()
    dev_fast = np.abs(got[..., 0] - want[..., 0]).max()
    dev_witness = np.abs(witness[..., 0] - want[..., 0]).max()
    assert np.quantile((want[..., 0] - st[..., 0])[2:-2, 2:-2], 0.01) > 1.0e-5   # it rained on (nearly) every cell
    assert dev_witness > 1e-9                                               # the premise: even the witness is off
    assert dev_fast <= 3.0 * dev_witness and dev_fast <= 1e-6
    assert sg["batch_successful"] == so["batch_successful"] == iters
    assert abs(sg["timestep"] - so["timestep"]) <= 1e-9 * so["timestep"]
    assert int(((got[..., 0] - bed) > 1e-10).sum()) == int(((want[..., 0] - bed) > 1e-10).sum())
    vol_o, vol_g = (want[..., 0] - bed).sum(), (got[..., 0] - bed).sum()
    assert abs(vol_g - vol_o) <= 1e-10 * vol_o


def test_march_runs_really_wrap():
    """The sizes above put >= 128 units on every CTA of the marching kernels (the condition for march_runs >= 2)."""
    def per_cta(rows, cols, use, ctas_per_sm):
        nstrips = -(-cols // use)
        ngroups = -(-nstrips // 4)
        return ngroups * rows // (ctas_per_sm * 148)
    assert per_cta(3072, 4096, 30, 4) >= 128      # MH / Godunov-march fp64
    assert per_cta(3072, 4096, 28, 6) >= 128      # MH / Godunov-march fp32
    assert per_cta(4096, 4096, 30, 6) >= 128      # inertial fp64
    assert per_cta(4096, 4096, 28, 8) >= 128      # inertial fp32
    assert (1024 // 8) * (1536 // 32) > 9 * 148   # Godunov tiles
    assert per_cta(4608, 3072, 30, 4) >= 128 and per_cta(4608, 3072, 30, 6) >= 128   # the wet/dry terrain cases


def test_configs3_combination(ex):
    """BASELINE configs[3] in small: partial-inertial fp32 + spatially gridded rain + point volume sources
    (CLSchemeInertial.clc:27-163 with bdy_Gridded CLBoundaries.clc:186-246 and bdy_Cell's volume branch :85-93)."""
    rows = cols = 512
    w = crop("radar16384", rows, cols)
    w["boundaries"] = None
    cfg = bench.cfg_for(w, rows, cols)
    bed, st, man = bench.make_inputs(w, rows, cols, np.float32)
    rng = np.random.default_rng(5)
    frames = rng.uniform(0.0, 80.0, size=(6, rows // 32 + 1, cols // 32 + 1))
    pts = sorted(set(int(y) * cols + int(x) for y, x in rng.integers(2, rows - 2, size=(64, 2))))
    series = [[60.0 * i, 0.0, q, 0.0] for i, q in enumerate([0.0, 2.0, 1.0, 0.0, 0.0, 0.0])]
    orc = cpu_sim.CpuSim("oracle", cfg)
    gpu = hx.CudaScheme(ex, cfg)
    iters = 150
    for sim in (orc, gpu):
        sim.upload(st, bed, man)
        sim.add_gridded(hc.GRIDDED_RAIN_INTENSITY, 120.0, 32.0, 0.0, 0.0, frames)
        sim.add_cell(hc.DEPTH_IGNORE, hc.DISCHARGE_IS_VOLUME, pts, series)
        sim.set_target(1.0e7)
        sim.set_clock(30.0, cfg.initial_dt, 0.97)
        sim.iterate(iters)
    check(cfg, orc, gpu, bed, st, iters)
    # the sources really fired: water appeared at the surcharging cells and rain fell
    out = gpu.download()
    assert (out[..., 0].astype(np.float64) - bed).sum() > (st[..., 0].astype(np.float64) - bed).sum()
    gpu.close()
    orc.close()


@pytest.mark.parametrize("workload", ["dambreak4096-inertial", "dambreak4096-inertial-f32", "radar16384", "dambreak4096-mh",
                                      "dambreak4096-mh-f32", "river32768", "pluvial16384"])
def test_wide_and_narrow_marching_kernels_agree(ex, workload):
    """The two-columns-per-lane ("wide") kernels evaluate, component by component, the expression tree of the
    one-column kernels they replace; only where the compiler contracts a product and a sum differently can the last
    bit differ.  Ragged sizes (columns not a multiple of 60, 30 or 28), boundaries attached, 25 iterations."""
    rows, cols, iters = 1500, 1999, 25
    w = crop(workload, rows, cols)
    cfg = bench.cfg_for(w, rows, cols)
    dtype = np.float64 if cfg.precision == "double" else np.float32
    bed, st, man = bench.make_inputs(w, rows, cols, dtype)
    out = []
    for options in (WIDE, NARROW):
        sim = hx.CudaScheme(ex, cfg, options=options)
        sim.upload(st, bed, man)
        bench.attach_boundaries(sim, w, cols, rows)
        sim.set_target(1.0e7)
        if w.get("boundaries"):
            sim.set_clock(100.0, cfg.initial_dt, 0.97)
        sim.iterate(iters)
        out.append((sim.download_both(), sim.stats()))
        sim.close()
    (wa, wb), sw = out[0]
    (na, nb), sn = out[1]
    assert sw["batch_successful"] == sn["batch_successful"] == iters
    # inertial: the same expression tree; MUSCL-Hancock: the wide kernel fuses h +- s/2 and the flux sums differently
    if cfg.scheme == hc.SCHEME_INERTIAL:
        tol = 1e-12 if cfg.precision == "double" else 1e-5
    else:
        tol = TOL[cfg.precision]
    assert abs(sw["time"] - sn["time"]) <= tol * max(1.0, sn["time"])
    for got, want in ((wa, na), (wb, nb)):
        assert np.isfinite(got).all()
        assert np.abs(got[..., :2] - want[..., :2]).max() <= tol
        assert np.abs(got[..., 2:] - want[..., 2:]).max() <= 100 * tol
