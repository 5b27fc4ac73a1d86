"""CPU restatement of the output-raster derivation, for the parity tests only.

TEST INFRASTRUCTURE (see oracle/hpo_api.h): follows CRasterDataset::domainToRaster,
src/Datasets/CRasterDataset.cpp:180-280 of the reference, value by value, in float64 like the
reference's host loop (numpy evaluates every operation separately: no FMA contraction).

Pinned to the reference: the routine itself is a GDAL writer and cannot be built here, but its only
arithmetic -- the per-cell `switch( ucValue )` -- is cut out of the reference source where it lies and
compiled between stand-ins for pDomain / pBand (oracle/build_ref.py:build_raster,
oracle/ref_shim/raster_shim.h).  This restatement is bit-identical to it on adversarial cells
(tests/test_raster_outputs.py, live and through tests/golden/raster_values.npz).
"""
import numpy as np

# model::rasterDatasets::dataValues, src/Datasets/CRasterDataset.h:33-46
DEPTH, FSL, VELOCITY_X, VELOCITY_Y, DISCHARGE_X, DISCHARGE_Y, MAX_DEPTH, MAX_FSL, FROUDE = 1, 2, 3, 4, 5, 6, 9, 10, 11


def derive_raster(value, states, bed, resolution, nodata=-9999.0):
    """states: rows x cols x 4 {eta, eta_max, qx, qy}, bed: rows x cols, row 0 = SOUTH.
    Returns rows x cols float64 with row 0 = NORTH (the RasterIO call writes row `rows - iRow - 1`, :270-280)."""
    st = np.asarray(states, dtype=np.float64)
    z = np.asarray(bed, dtype=np.float64)
    eta, emax, qx, qy = st[..., 0], st[..., 1], st[..., 2], st[..., 3]
    depth = eta - z
    nd = np.float64(nodata)
    with np.errstate(divide="ignore", invalid="ignore"):
        if value == MAX_FSL:                                   # :187-196
            out = np.where((emax < z + 1e-8) | (z > 9999.0), nd, emax)
        elif value == FSL:                                     # :197-206
            out = np.where((eta < z + 1e-8) | (z > 9999.0), nd, eta)
        elif value == MAX_DEPTH:                               # :207-213
            d = np.maximum(0.0, emax - z)
            out = np.where((d < 1e-8) | (d <= -9990.0) | (d >= 9999.0), nd, d)
        elif value == DEPTH:                                   # :214-220
            d = np.maximum(0.0, depth)
            out = np.where(d < 1e-8, nd, d)
        elif value == DISCHARGE_X:                             # :221-226
            out = qx * resolution
        elif value == DISCHARGE_Y:                             # :227-232
            out = qy * resolution
        elif value == VELOCITY_X:                              # :233-242
            out = np.where(depth > 1e-8, qx / depth, nd)
        elif value == VELOCITY_Y:                              # :243-252
            out = np.where(depth > 1e-8, qy / depth, nd)
        elif value == FROUDE:                                  # :253-266
            u, v = qx / depth, qy / depth
            out = np.where(depth > 1e-8, np.sqrt(u * u + v * v) / np.sqrt(9.81 * depth), nd)
        else:                                                  # :183, the row is pre-filled with -9999
            out = np.full(z.shape, nd)
    return np.ascontiguousarray(out[::-1])
