"""Output rasters: the device derivation against the oracle, and the oracle against
hand-computed cases of CRasterDataset::domainToRaster (src/Datasets/CRasterDataset.cpp:180-280)."""
import os

import numpy as np
import pytest

from hipims_ocl_b200 import config as hc
from oracle import raster_oracle as ro
from tests.helpers import dtype_of, make_cfg, scenario

ALL_VALUES = ["depth", "fsl", "velocityx", "velocityy", "dischargex", "dischargey", "maxdepth", "maxfsl", "froude"]


def test_oracle_matches_hand_computed_cases():
    bed = np.array([[1.0, 2.0, 10000.0], [0.5, 0.5, 0.5]])
    st = np.zeros((2, 3, 4))
    st[0, 0] = [1.5, 1.75, 0.25, -0.5]          # wet: depth 0.5
    st[0, 1] = [2.0, 2.0, 0.0, 0.0]             # dry
    st[0, 2] = [10000.0, 10000.0, 0.0, 0.0]     # bed above 9999: no data for levels
    st[1, 0] = [0.5 + 5e-9, 0.5, 1.0, 1.0]      # thinner than 1e-8: no data
    st[1, 1] = [2.5, 3.0, 2.0, 0.0]
    st[1, 2] = [0.5, -9999.0, 0.0, 0.0]         # disabled cell
    nd = -9999.0
    depth = ro.derive_raster(ro.DEPTH, st, bed, 2.0)
    assert depth.tolist() == [[nd, 2.0, nd], [0.5, nd, nd]]          # row 0 is the NORTH row
    assert ro.derive_raster(ro.FSL, st, bed, 2.0).tolist() == [[nd, 2.5, nd], [1.5, nd, nd]]
    assert ro.derive_raster(ro.MAX_DEPTH, st, bed, 2.0).tolist() == [[nd, 2.5, nd], [0.75, nd, nd]]
    assert ro.derive_raster(ro.MAX_FSL, st, bed, 2.0).tolist() == [[nd, 3.0, nd], [1.75, nd, nd]]
    assert ro.derive_raster(ro.DISCHARGE_X, st, bed, 2.0).tolist() == [[2.0, 4.0, 0.0], [0.5, 0.0, 0.0]]
    assert ro.derive_raster(ro.VELOCITY_X, st, bed, 2.0).tolist() == [[nd, 1.0, nd], [0.5, nd, nd]]
    assert ro.derive_raster(ro.VELOCITY_Y, st, bed, 2.0).tolist() == [[nd, 0.0, nd], [-1.0, nd, nd]]
    fr = ro.derive_raster(ro.FROUDE, st, bed, 2.0)
    assert fr[1, 0] == np.sqrt(0.5 * 0.5 + 1.0) / np.sqrt(9.81 * 0.5) and fr[0, 1] == 1.0 / np.sqrt(9.81 * 2.0)
    assert fr[0, 0] == nd and fr[1, 1] == nd
    assert ro.derive_raster(0, st, bed, 2.0).tolist() == [[nd] * 3, [nd] * 3]      # codes without a case stay no-data
    assert set(hc.RASTER_VALUES.values()) == {ro.DEPTH, ro.FSL, ro.VELOCITY_X, ro.VELOCITY_Y, ro.DISCHARGE_X, ro.DISCHARGE_Y,
                                              ro.MAX_DEPTH, ro.MAX_FSL, ro.FROUDE}


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "raster_values.npz")


def test_oracle_matches_the_reference_switch_golden():
    """tests/golden/raster_values.npz holds what the reference's own `switch( ucValue )` of CRasterDataset::domainToRaster
    (compiled from /root/reference by oracle/build_ref.py:build_raster) returns for 960 adversarial cells."""
    z = np.load(GOLDEN)
    for code in range(12):
        np.testing.assert_array_equal(ro.derive_raster(code, z["states"], z["bed"], float(z["resolution"])), z["value_%d" % code],
                                      err_msg="value code %d" % code)


def test_oracle_matches_the_reference_switch_live():
    """Same check against the reference source itself, on fresh random cells (skipped where /root/reference is absent)."""
    import ctypes as C
    from oracle import build_ref
    if not build_ref.reference_available():
        pytest.skip("reference tree not present")
    lib = C.CDLL(build_ref.build_raster())
    lib.ref_raster_value.restype = C.c_double
    lib.ref_raster_value.argtypes = [C.c_ubyte, C.POINTER(C.c_double), C.c_double, C.c_double]
    rng = np.random.default_rng(99)
    n = 600
    bed = rng.uniform(-10.0, 100.0, size=(1, n))
    depth = np.where(rng.random((1, n)) < 0.3, 0.0, 10.0 ** rng.uniform(-10.0, 1.0, size=(1, n)))
    st = np.zeros((1, n, 4))
    st[..., 0] = bed + depth
    st[..., 1] = st[..., 0] + rng.uniform(0.0, 0.5, size=(1, n)) * (rng.random((1, n)) < 0.5)
    st[..., 2:] = rng.normal(size=(1, n, 2))
    for code in range(12):
        want = np.array([[lib.ref_raster_value(code, np.ascontiguousarray(st[0, i]).ctypes.data_as(C.POINTER(C.c_double)), float(bed[0, i]), 3.0)
                          for i in range(n)]])
        np.testing.assert_array_equal(ro.derive_raster(code, st, bed, 3.0), want, err_msg="value code %d" % code)


@pytest.mark.gpu
def test_device_rasters_match_the_reference_switch_golden():
    from hipims_ocl_b200 import executor as hx
    z = np.load(GOLDEN)
    st, bed = z["states"], z["bed"]
    rows, cols = bed.shape
    cfg = make_cfg("godunov", "double", rows, cols, delta=float(z["resolution"]))
    ex = hx.Executor(0)
    sim = hx.CudaScheme(ex, cfg)
    sim.upload(st, bed, np.full((rows, cols), 0.03))
    for code in range(12):
        np.testing.assert_array_equal(sim.derive_raster(code), z["value_%d" % code], err_msg="value code %d" % code)
    sim.close()
    ex.close()


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["double", "single"])
@pytest.mark.parametrize("scheme", ["godunov", "muscl-hancock"])
def test_device_rasters_are_bit_identical_to_the_oracle(scheme, precision):
    from hipims_ocl_b200 import executor as hx
    rows, cols = 53, 71
    cfg = make_cfg(scheme, precision, rows, cols, delta=2.0)
    bed, st, man = scenario("wetdry", rows, cols, dtype_of(precision), seed=5)
    st[7, 9, 1] = -9999.0                                   # a disabled cell
    ex = hx.Executor(0)
    sim = hx.CudaScheme(ex, cfg)
    sim.upload(st, bed, man)
    sim.set_target(1e6)
    for iters in (0, 25):                                    # initial state, then an odd count so the ping-pong has swapped
        sim.iterate(iters)
        state = sim.download()
        for name in ALL_VALUES:
            got = sim.derive_raster(name)
            want = ro.derive_raster(hc.RASTER_VALUES[name], state, bed, cfg.delta)
            np.testing.assert_array_equal(got, want, err_msg=name)
        assert (sim.derive_raster("depth") != -9999.0).any()
    assert (sim.derive_raster(0) == -9999.0).all()
    sim.close()
    ex.close()


@pytest.mark.gpu
def test_device_raster_in_chunks_at_scale():
    """2048 x 2048 (larger than one staging chunk would be at 16384^2 is not needed: the chunk loop is exercised by
    the row arithmetic) -- depth raster equals eta - bed flipped, no-data where dry."""
    from hipims_ocl_b200 import executor as hx
    n = 2048
    cfg = make_cfg("godunov", "double", n, n)
    bed, st, man = scenario("dambreak-dry", n, n, np.float64)
    ex = hx.Executor(0)
    sim = hx.CudaScheme(ex, cfg)
    sim.upload(st, bed, man)
    got = sim.derive_raster("depth")
    d = (st[..., 0] - bed)[::-1]
    np.testing.assert_array_equal(got, np.where(d < 1e-8, -9999.0, d))
    sim.close()
    ex.close()
