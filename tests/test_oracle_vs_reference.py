"""The oracle (oracle/hipims_oracle.cpp) must be BIT-IDENTICAL to the reference's own kernel
sources compiled through oracle/ref_shim (both -ffp-contract=off).  CPU only; skipped when
neither /root/reference nor a prebuilt oracle/_ref library is available."""
import numpy as np
import pytest

from hipims_ocl_b200 import config as hc
from oracle import cpu_sim
from tests.helpers import add_standard_boundaries, dtype_of, make_cfg, scenario


def _pair(cfg):
    if not cpu_sim.ref_available(cfg):
        pytest.skip("reference tree / prebuilt reference library not available")
    return cpu_sim.CpuSim("oracle", cfg), cpu_sim.CpuSim("ref", cfg)


def _same(a, b):
    # bit-identical apart from the sign of zero
    np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("precision", ["double", "single"])
@pytest.mark.parametrize("scheme", ["godunov", "inertial"])
def test_single_step_kernel(scheme, precision):
    rows, cols = 61, 83
    cfg = make_cfg(scheme, precision, rows, cols)
    orc, ref = _pair(cfg)
    dt = dtype_of(precision)
    for seed in (1, 2, 3):
        bed, st, man = scenario("wetdry", rows, cols, dt, seed)
        for step_dt in (0.05, 0.0, -0.01):
            d_o = np.full_like(st, -123.0)
            d_r = np.full_like(st, -123.0)
            orc.k_step(step_dt, bed, st, d_o, man)
            ref.k_step(step_dt, bed, st, d_r, man)
            _same(d_o, d_r)
            if step_dt > 0:
                assert (d_o != -123.0).any() and (d_o == -123.0).any()  # ring (and dry stencils) untouched
    assert abs(orc.k_reduce(st, bed) - ref.k_reduce(st, bed)) == 0.0


def test_godunov_without_friction():
    rows, cols = 40, 52
    cfg = make_cfg("godunov", "double", rows, cols, friction=False)
    orc, ref = _pair(cfg)
    bed, st, man = scenario("wetdry", rows, cols, np.float64, 5)
    d_o, d_r = np.full_like(st, -7.0), np.full_like(st, -7.0)
    orc.k_step(0.04, bed, st, d_o, man)
    ref.k_step(0.04, bed, st, d_r, man)
    _same(d_o, d_r)


@pytest.mark.parametrize("precision", ["double", "single"])
def test_muscl_hancock_kernels(precision):
    rows, cols = 57, 66
    cfg = make_cfg("muscl-hancock", precision, rows, cols)
    orc, ref = _pair(cfg)
    dt = dtype_of(precision)
    for seed in (4, 5):
        bed, st, man = scenario("wetdry", rows, cols, dt, seed)
        f_o = [np.full_like(st, -5.0) for _ in range(4)]
        f_r = [np.full_like(st, -5.0) for _ in range(4)]
        orc.k_mch_1st(0.03, bed, st, f_o)
        ref.k_mch_1st(0.03, bed, st, f_r)
        for a, b in zip(f_o, f_r):
            _same(a, b)
        s_o, s_r = st.copy(), st.copy()
        orc.k_mch_2nd(0.03, s_o, bed, man, f_o)
        ref.k_mch_2nd(0.03, s_r, bed, man, f_r)
        _same(s_o, s_r)
        assert (s_o != st).any()


CASES = [
    # scheme, precision, scenario, boundaries, n, iterations, extra config
    ("godunov", "double", "dambreak", "none", 64, 60, {}),
    ("godunov", "double", "dambreak-dry", "none", 64, 60, {}),
    ("godunov", "single", "dambreak", "none", 64, 60, {}),
    ("godunov", "double", "pluvial", "rain+loss", 50, 400, {"delta": 2.0}),
    ("godunov", "double", "pluvial-wet", "gridded", 48, 150, {}),
    ("godunov", "double", "valley", "cells", 48, 150, {}),
    ("godunov", "double", "dambreak", "none", 48, 40, {"dynamic": False, "fixed_dt": 0.01}),
    ("godunov", "double", "wetdry", "rain", 45, 80, {"quirks": 0}),
    ("inertial", "double", "pluvial-wet", "rain", 48, 200, {}),
    ("inertial", "single", "valley", "cells", 48, 150, {}),
    ("muscl-hancock", "double", "dambreak", "none", 64, 50, {}),
    ("muscl-hancock", "double", "dambreak-dry", "none", 64, 50, {}),
    ("muscl-hancock", "single", "dambreak", "none", 48, 50, {}),
    ("muscl-hancock", "double", "valley", "cells", 48, 120, {}),
    ("muscl-hancock", "double", "pluvial-wet", "rain", 48, 200, {"quirks": hc.QUIRKS_REFERENCE | hc.QUIRK_MH_NO_BOUNDARIES}),
]


@pytest.mark.parametrize("scheme,precision,scen,bdy,n,iters,extra", CASES)
def test_multi_step_runs(scheme, precision, scen, bdy, n, iters, extra):
    cfg = make_cfg(scheme, precision, n, n, **extra)
    orc, ref = _pair(cfg)
    bed, st, man = scenario(scen, n, n, dtype_of(precision))
    for sim in (orc, ref):
        sim.upload(st, bed, man)
        add_standard_boundaries(sim, cfg, bdy)
        sim.set_target(1.0e6)
    done = 0
    for chunk in (1, 2, 7, iters - 10):
        orc.iterate(chunk)
        ref.iterate(chunk)
        done += chunk
        assert orc.stats() == ref.stats(), "clock diverged after %d iterations" % done
        a_o, b_o = orc.download_both()
        a_r, b_r = ref.download_both()
        _same(a_o, a_r)
        _same(b_o, b_r)
    s = orc.stats()
    assert s["batch_successful"] == iters and s["time"] > 0.0
    assert np.isfinite(orc.download()).all()


def test_sync_point_suspension_and_update():
    """dt goes negative at the sync time (CLDynamicTimestep.clc:118-124), iterations are then
    skipped; a new target plus tst_UpdateTimestep resumes (CSchemeGodunov.cpp:1164-1210)."""
    n = 40
    cfg = make_cfg("godunov", "double", n, n)
    orc, ref = _pair(cfg)
    bed, st, man = scenario("dambreak", n, n, np.float64)
    for sim in (orc, ref):
        sim.upload(st, bed, man)
        sim.set_target(0.25)
        sim.iterate(30)
    so, sr = orc.stats(), ref.stats()
    assert so == sr
    assert so["time"] == 0.25 and so["timestep"] < 0.0 and so["batch_skipped"] > 0
    for sim in (orc, ref):
        sim.set_target(0.5)
        sim.update_timestep()
        sim.reset_counters()
        sim.iterate(10)
    assert orc.stats() == ref.stats()
    _same(orc.download(), ref.download())
    assert orc.stats()["time"] > 0.25


def test_godunov_dt0_keep_rule():
    """The two Godunov kernels of the reference differ in one observable rule: with a timestep <= 0 gts_cacheDisabled
    copies the source cell into the destination (CLSchemeGodunov.clc:201-206), gts_cacheEnabled returns before any write
    (:477-478).  The oracle restates both (HPO_QUIRK_GODUNOV_DT0_KEEP selects the second); both are pinned to the compiled
    reference kernels, the second in test_local_memory_kernels_of_the_reference.  Here: what the rule does to a run."""
    n = 40
    bed, st, man = scenario("dambreak", n, n, np.float64)
    out = {}
    for quirk in (0, hc.QUIRK_GODUNOV_DT0_KEEP):
        cfg = make_cfg("godunov", "double", n, n)
        cfg.quirks |= quirk
        orc = cpu_sim.CpuSim("oracle", cfg)
        orc.upload(st, bed, man)
        orc.set_target(0.25)
        orc.iterate(30)
        assert orc.stats()["timestep"] < 0.0 and orc.stats()["batch_skipped"] > 2
        before = orc.download_both()
        orc.iterate(1)
        out[quirk] = (before, orc.download_both(), orc.stats())
        orc.close()
    (a0, b0), (a1, b1), s_copy = out[0]
    np.testing.assert_array_equal(a1, b1)                      # copied through: both buffers hold the state at the target
    (a0, b0), (a1, b1), s_keep = out[hc.QUIRK_GODUNOV_DT0_KEEP]
    np.testing.assert_array_equal(a1, a0)                      # nothing written ...
    np.testing.assert_array_equal(b1, b0)
    assert not np.array_equal(a1, b1)                          # ... so the older buffer stays one step behind
    assert s_keep["time"] == s_copy["time"] == 0.25


@pytest.mark.parametrize("precision", ["double", "single"])
@pytest.mark.parametrize("scheme", ["godunov", "inertial"])
def test_local_memory_kernels_of_the_reference(scheme, precision):
    """gts_cacheEnabled / ine_cacheEnabled -- the reference's step kernels with the tile staged in work-group local memory
    -- run through the shim a work-group at a time (16 x 16 items, groups overlapping by two cells, one barrier;
    oracle/ref_shim/ref_unit.inc: ndrange2_tiles).  They are alternative implementations of the same maths: with a
    positive timestep bit-identical to the oracle.  With a timestep <= 0 both copy disabled cells through and leave every
    other cell untouched (CLSchemeGodunov.clc:449-454, 477-478; CLSchemeInertial.clc:215-241), where gts_cacheDisabled
    copies the whole interior (:201-206; the oracle restates the other rule under HPO_QUIRK_GODUNOV_DT0_KEEP) and
    ine_cacheDisabled returns before even the disabled-cell copy (CLSchemeInertial.clc:60-61 -- no switch for that one: a
    disabled cell holds the same values in both ping-pong buffers anyway).  Sizes that are not a multiple of the 14 cells
    a group updates; the destination starts as the source with the interior levels moved, so that "left untouched" and
    "copied through" are both visible."""
    rows, cols = 61, 83
    cfg = make_cfg(scheme, precision, rows, cols)
    cfg_keep = make_cfg(scheme, precision, rows, cols)
    cfg_keep.quirks |= hc.QUIRK_GODUNOV_DT0_KEEP
    orc, ref = _pair(cfg)
    orc_keep = cpu_sim.CpuSim("oracle", cfg_keep)
    dt = dtype_of(precision)
    for seed in (1, 2):
        bed, st, man = scenario("wetdry", rows, cols, dt, seed)
        off = (st[..., 1] <= -9999.0) | (st[..., 0] == -9999.0)
        inner = np.zeros(off.shape, bool)
        inner[1:-1, 1:-1] = True
        assert (off & inner).any()                                        # disabled cells inside the domain
        start = st.copy()
        start[1:-1, 1:-1, 0] += 1.0
        for step_dt in (0.05, 0.0, -0.01):
            d_ref, d_orc, d_keep = start.copy(), start.copy(), start.copy()
            ref.k_step_cached(step_dt, bed, st, d_ref, man)
            orc.k_step(step_dt, bed, st, d_orc, man)
            orc_keep.k_step(step_dt, bed, st, d_keep, man)
            if step_dt > 0:
                _same(d_orc, d_ref)
                _same(d_keep, d_ref)
                assert (d_ref != start).any()
                continue
            expected = start.copy()
            expected[off & inner] = st[off & inner]                       # disabled cells copied, the rest untouched
            _same(d_ref, expected)
            if scheme == "godunov":
                _same(d_keep, d_ref)                                      # the quirk IS the local-memory kernel's rule
                assert not np.array_equal(d_orc, d_ref)                   # gts_cacheDisabled copied the interior through
            else:
                _same(d_orc, start)                                       # ine_cacheDisabled wrote nothing at all
    orc_keep.close()


@pytest.mark.parametrize("precision", ["double", "single"])
def test_local_memory_predictor_of_the_reference(precision):
    """mch_1st_cachePrediction (the MUSCL-Hancock predictor with its tile in work-group local memory,
    CLSchemeMUSCLHancock.clc:158-) through the same work-group emulation: the four face buffers are bit-identical to the
    oracle's (and so to mch_1st_cacheNone's), for a positive and for a non-positive timestep -- except next to DISABLED
    cells: the local tile carries the bed elevation in the slot of eta_max (:199-200), so the variant's "is a neighbour
    disabled" test (:246-251, and the one inside mch_1st) reads a bed where mch_1st_cacheNone reads eta_max, and keeps the
    second-order predictor where the default kernel falls back to the cell's own state.  One more reason (with SURVEY Q3)
    why the oracle follows the default pair mch_1st_cacheNone + mch_2nd_cacheNone."""
    from hipims_ocl_b200 import scenarios as sc
    rows, cols = 57, 66
    cfg = make_cfg("muscl-hancock", precision, rows, cols)
    orc, ref = _pair(cfg)
    dt = dtype_of(precision)
    for seed in (4, 5):
        for disabled in (False, True):
            bed, st, man = sc.random_wet_dry(rows, cols, seed, dtype=dt, disabled=disabled)
            off = st[..., 1] <= -9999.0
            near = np.zeros_like(off)
            near[1:, :] |= off[:-1, :]; near[:-1, :] |= off[1:, :]; near[:, 1:] |= off[:, :-1]; near[:, :-1] |= off[:, 1:]
            for step_dt in (0.03, 0.0):
                f_o = [np.full_like(st, -5.0) for _ in range(4)]
                f_r = [np.full_like(st, -5.0) for _ in range(4)]
                orc.k_mch_1st(step_dt, bed, st, f_o)
                ref.k_mch_1st_cached(step_dt, bed, st, f_r)
                for a, b in zip(f_o, f_r):
                    differs = (a != b).any(axis=2)
                    if disabled:
                        assert not (differs & ~(near | off)).any()       # only beside (or on) disabled cells
                    else:
                        assert not differs.any()
                if step_dt > 0:
                    assert (f_o[0] != -5.0).any()
