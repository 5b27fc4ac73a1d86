// raster_shim.h -- stand-ins that let the per-cell switch of CRasterDataset::domainToRaster
// (src/Datasets/CRasterDataset.cpp of the reference) compile on its own.  TEST INFRASTRUCTURE ONLY
// (see oracle/build_ref.py); the switch text itself is read from the reference tree at build time.
#pragma once
#include <algorithm>
#include <cmath>
using std::max;
using std::min;

namespace model {
namespace rasterDatasets { namespace dataValues { enum dataValues {      // src/Datasets/CRasterDataset.h:33-46
    kBedElevation = 0, kDepth = 1, kFreeSurfaceLevel = 2, kVelocityX = 3, kVelocityY = 4, kDischargeX = 5, kDischargeY = 6,
    kManningCoefficient = 7, kDisabledCells = 8, kMaxDepth = 9, kMaxFSL = 10, kFroudeNumber = 11 }; } }
namespace domainValueIndices { enum domainValueIndices {                 // src/Domain/CDomain.h:28-33
    kValueFreeSurfaceLevel = 0, kValueMaxFreeSurfaceLevel = 1, kValueDischargeX = 2, kValueDischargeY = 3 }; }
}

struct RasterShimDomain {                 // CDomain::getStateValue / getBedElevation on one cell
    const double* state; double bed;
    double getStateValue(unsigned long, int index) const { return state[index]; }
    double getBedElevation(unsigned long) const { return bed; }
};
struct RasterShimBand { double GetNoDataValue() const { return -9999.0; } };   // pBand->SetNoDataValue( -9999.0 )

// the locals of domainToRaster the switch uses (CRasterDataset.cpp:115-124, 176-183)
#define RASTER_SHIM_PROLOGUE                                                   \
    RasterShimDomain shimDomain{state4, bed}; RasterShimDomain* pDomain = &shimDomain; \
    RasterShimBand shimBand; RasterShimBand* pBand = &shimBand;               \
    double dRowStore[1]; double* dRow = dRowStore; unsigned long iCol = 0, ulCellID = 0; \
    double dDepth = 0, dVelocityX = 0, dVelocityY = 0;                        \
    dRow[iCol] = -9999.0;
