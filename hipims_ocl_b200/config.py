"""Scheme configuration shared by the host-side mirror and the tests.

Field names follow the reference's XML scheme parameters
(src/Schemes/CSchemeGodunov.cpp:128-334, src/Schemes/CScheme.cpp:49-109) and the JIT
constants they end up in (src/Schemes/CSchemeGodunov.cpp:667-783).
"""
from dataclasses import dataclass, replace

SCHEME_GODUNOV = "godunov"
SCHEME_MUSCL_HANCOCK = "muscl-hancock"
SCHEME_INERTIAL = "inertial"

# Reference quirks that can be reproduced or switched off.
QUIRK_REDUCE_BUFFER_A = 1    # Q1: the CFL reduction always reads buffer "Cell states"
QUIRK_BDY_COVERAGE = 2       # Q6: bdy_Uniform/bdy_Gridded cover floor(n/8)*8 cells per axis
QUIRK_MH_NO_BOUNDARIES = 4   # Q4: MUSCL-Hancock never applies boundary kernels
QUIRK_GODUNOV_DT0_KEEP = 8   # gts_cacheEnabled's rule: a Godunov step with timestep <= 0 writes nothing (default: copies through)
QUIRKS_REFERENCE = QUIRK_REDUCE_BUFFER_A | QUIRK_BDY_COVERAGE

# src/Boundaries/CLBoundaries.clh:31-52
DEPTH_IGNORE, DEPTH_IS_FSL, DEPTH_IS_DEPTH, DEPTH_IS_CRITICAL = 0, 1, 2, 3
DISCHARGE_IGNORE, DISCHARGE_IS_DISCHARGE, DISCHARGE_IS_VELOCITY, DISCHARGE_IS_VOLUME = 0, 1, 2, 3
UNIFORM_RAIN_INTENSITY, UNIFORM_LOSS_RATE = 0, 1
GRIDDED_RAIN_INTENSITY, GRIDDED_MASS_FLUX = 0, 2

# Output raster values: model::rasterDatasets::dataValues (src/Datasets/CRasterDataset.h:33-46), keyed by the
# names CDomain::getDataValueCode accepts in <dataTarget value="..."> (src/Domain/CDomain.cpp:464-500)
RASTER_DEPTH, RASTER_FSL, RASTER_VELOCITY_X, RASTER_VELOCITY_Y, RASTER_DISCHARGE_X, RASTER_DISCHARGE_Y = 1, 2, 3, 4, 5, 6
RASTER_MAX_DEPTH, RASTER_MAX_FSL, RASTER_FROUDE = 9, 10, 11
RASTER_VALUES = {"depth": 1, "fsl": 2, "velocityx": 3, "velocityy": 4, "dischargex": 5, "dischargey": 6, "maxdepth": 9,
                 "maxfsl": 10, "froude": 11}


@dataclass
class SchemeConfig:
    scheme: str = SCHEME_GODUNOV          # <scheme name="...">
    precision: str = "double"             # floatingPointPrecision
    cols: int = 0
    rows: int = 0
    delta: float = 1.0                    # cell resolution (DOMAIN_DELTAX == DOMAIN_DELTAY)
    courant: float = 0.5                  # courantNumber
    dry_threshold: float = 1e-10          # dryThreshold (VERY_SMALL)
    end_time: float = 3600.0              # simulation duration (SCHEME_ENDTIME)
    dynamic: bool = True                  # timestepMode auto|cfl vs fixed
    fixed_dt: float = 0.0                 # timestepFixed
    initial_dt: float = 0.001             # timestepInitial (src/Schemes/CScheme.cpp:49)
    friction: bool = True                 # frictionEffects
    quirks: int = QUIRKS_REFERENCE

    @property
    def real_bytes(self):
        return 8 if self.precision == "double" else 4

    @property
    def cells(self):
        return self.cols * self.rows

    def with_(self, **kw):
        return replace(self, **kw)
