#!/usr/bin/env python3
"""Build the reference's own kernel sources into CPU libraries under oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  The reference host program cannot be built in this image (no
Boost/GDAL/OpenCL headers), but its kernel sources are plain OpenCL C
and compile with g++ through oracle/ref_shim/cl_shim.h.  This script

  1. reads the .clh/.clc files WHERE THEY LIE under /root/reference (never copied into the
     repository), in the order the reference's prepareCode() concatenates them
     (src/Schemes/CSchemeGodunov.cpp:483-504, CSchemeMUSCLHancock.cpp:324-345,
     CSchemeInertial.cpp:204-220), after the universal header the program wrapper always
     prepends (src/OpenCL/Executors/COCLProgram.cpp);
  2. applies two mechanical rewrites (vector-constructor syntax, work-group attributes);
  3. writes the translation unit to a temp directory and compiles it with
     g++ -O2 -ffp-contract=off -fopenmp into oracle/_ref/ref_<variant>.so.

Variants: scheme in {godunov, mh, inertial} x precision {f64, f32} x timestep {dyn, fix}
x friction {fric, nofric}.  `python oracle/build_ref.py` builds the set the tests use.
"""
import os
import re
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("HIPIMS_REFERENCE", "/root/reference") + "/src"
OUT = os.path.join(HERE, "_ref")

_COMMON_H = ["Domain/Cartesian/CLDomainCartesian.clh", "Schemes/CLFriction.clh"]
_COMMON_C = ["Domain/Cartesian/CLDomainCartesian.clc", "Schemes/CLFriction.clc"]
PROGRAMS = {
    "godunov": (_COMMON_H + ["Solvers/CLSolverHLLC.clh", "Schemes/CLDynamicTimestep.clh", "Schemes/CLSchemeGodunov.clh",
                             "Boundaries/CLBoundaries.clh"],
                _COMMON_C + ["Solvers/CLSolverHLLC.clc", "Schemes/CLDynamicTimestep.clc", "Schemes/CLSchemeGodunov.clc",
                             "Boundaries/CLBoundaries.clc"], "SHIM_SCHEME_GODUNOV"),
    "mh": (_COMMON_H + ["Schemes/Limiters/CLSlopeLimiterMINMOD.clh", "Solvers/CLSolverHLLC.clh",
                        "Schemes/CLDynamicTimestep.clh", "Schemes/CLSchemeMUSCLHancock.clh",
                        "Boundaries/CLBoundaries.clh"],
           _COMMON_C + ["Schemes/Limiters/CLSlopeLimiterMINMOD.clc", "Solvers/CLSolverHLLC.clc",
                        "Schemes/CLDynamicTimestep.clc", "Schemes/CLSchemeMUSCLHancock.clc",
                        "Boundaries/CLBoundaries.clc"], "SHIM_SCHEME_MUSCL_HANCOCK"),
    "inertial": (_COMMON_H + ["Schemes/CLDynamicTimestep.clh", "Schemes/CLSchemeInertial.clh",
                              "Boundaries/CLBoundaries.clh"],
                 _COMMON_C + ["Schemes/CLDynamicTimestep.clc", "Schemes/CLSchemeInertial.clc",
                              "Boundaries/CLBoundaries.clc"], "SHIM_SCHEME_INERTIAL"),
}

DEFAULT_VARIANTS = [
    "godunov_f64_dyn_fric", "godunov_f64_dyn_nofric", "godunov_f64_fix_fric", "godunov_f32_dyn_fric",
    "mh_f64_dyn_fric", "mh_f32_dyn_fric", "inertial_f64_dyn_fric", "inertial_f32_dyn_fric",
]

_VEC_CTOR = re.compile(r"\(\s*(cl_double[248]|uint2|cl_uint2)\s*\)\s*\(")
_WG_ATTR = re.compile(r"__attribute__\s*\(\(\s*reqd_work_group_size\s*\([^)]*\)\s*\)\)")


def reference_available():
    return os.path.isdir(REF_SRC)


def _translation_unit(scheme):
    headers, sources, macro = PROGRAMS[scheme]
    parts = ['#include "cl_shim.h"\n']
    for rel in ["OpenCL/Executors/CLUniversalHeader.clh"] + headers + sources:
        with open(os.path.join(REF_SRC, rel), "r", encoding="latin-1") as f:
            text = f.read()
        text = _VEC_CTOR.sub(lambda mt: "shim_make<%s>(" % mt.group(1), text)
        text = _WG_ATTR.sub("", text)
        parts.append('#line 1 "%s"\n%s\n' % (rel, text))
    parts.append('#include "ref_unit.inc"\n')
    return "".join(parts), macro


def build_variant(name, force=False):
    scheme, prec, ts, fric = name.split("_")
    out = os.path.join(OUT, "ref_%s.so" % name)
    if os.path.exists(out) and not force:
        deps = [os.path.join(HERE, "ref_shim", "cl_shim.h"), os.path.join(HERE, "ref_shim", "ref_unit.inc"),
                os.path.join(HERE, "sim_driver.h"), os.path.join(HERE, "hpo_api.h"), os.path.abspath(__file__)]
        if all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
            return out
    os.makedirs(OUT, exist_ok=True)
    tu, macro = _translation_unit(scheme)
    flags = ["-D" + macro]
    if prec == "f32":
        flags += ["-DSHIM_REAL=float", "-fsingle-precision-constant"]
    if ts == "fix":
        flags.append("-DSHIM_TIMESTEP_FIXED")
    if fric == "fric":
        flags.append("-DSHIM_FRICTION")
    with tempfile.TemporaryDirectory(prefix="hpo_ref_") as tmp:
        src = os.path.join(tmp, "unit_%s.cpp" % name)
        with open(src, "w") as f:
            f.write(tu)
        cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fopenmp", "-fpermissive", "-w", "-shared", "-fPIC",
               "-I", os.path.join(HERE, "ref_shim"), "-I", HERE] + flags + [src, "-o", out]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("reference shim build failed for %s:\n%s" % (name, res.stderr[-4000:]))
    return out


# ---------------------------------------------------------------------------------------------
# Output rasters: the per-cell `switch( ucValue )` of CRasterDataset::domainToRaster
# (src/Datasets/CRasterDataset.cpp) is the only arithmetic of that GDAL writer.  Its text is cut out
# of the reference file where it lies and compiled between stand-ins for pDomain / pBand
# (oracle/ref_shim/raster_shim.h) into oracle/_ref/ref_raster.so: ref_raster_value(code, state4,
# bed, resolution) returns exactly what the reference would hand to RasterIO for one cell.
# ---------------------------------------------------------------------------------------------
RASTER_LIB = os.path.join(OUT, "ref_raster.so")


def _raster_switch_text():
    with open(os.path.join(REF_SRC, "Datasets/CRasterDataset.cpp"), "r", encoding="latin-1") as f:
        text = f.read()
    start = text.index("switch( ucValue )", text.index("CRasterDataset::domainToRaster"))
    depth, i = 0, text.index("{", start)
    for j in range(i, len(text)):
        depth += text[j] == "{"
        depth -= text[j] == "}"
        if depth == 0:
            return text[start:j + 1]
    raise RuntimeError("unbalanced switch in CRasterDataset.cpp")


def build_raster(force=False):
    shim = os.path.join(HERE, "ref_shim", "raster_shim.h")
    if os.path.exists(RASTER_LIB) and not force and \
            all(os.path.getmtime(RASTER_LIB) >= os.path.getmtime(d) for d in (shim, os.path.abspath(__file__))):
        return RASTER_LIB
    os.makedirs(OUT, exist_ok=True)
    tu = '#include "raster_shim.h"\nextern "C" double ref_raster_value(unsigned char ucValue, const double* state4, double bed, double dResolution) {\n' \
         '    RASTER_SHIM_PROLOGUE\n#line 1 "Datasets/CRasterDataset.cpp (switch)"\n' + _raster_switch_text() + '\n    return dRow[iCol];\n}\n'
    with tempfile.TemporaryDirectory(prefix="hpo_ref_") as tmp:
        src = os.path.join(tmp, "unit_raster.cpp")
        with open(src, "w") as f:
            f.write(tu)
        cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-w", "-shared", "-fPIC", "-I", os.path.join(HERE, "ref_shim"), src, "-o", RASTER_LIB]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("reference raster shim build failed:\n%s" % res.stderr[-4000:])
    return RASTER_LIB


# Util::round (src/util.cpp): the 4-decimal rounding every ingested DEM / depth value goes through.
UTIL_LIB = os.path.join(OUT, "ref_util.so")


def build_util(force=False):
    if os.path.exists(UTIL_LIB) and not force and os.path.getmtime(UTIL_LIB) >= os.path.getmtime(os.path.abspath(__file__)):
        return UTIL_LIB
    with open(os.path.join(REF_SRC, "util.cpp"), "r", encoding="latin-1") as f:
        text = f.read()
    start = text.index("double\tround( double dValue, unsigned char ucPlaces )")
    depth, i = 0, text.index("{", start)
    for j in range(i, len(text)):
        depth += text[j] == "{"
        depth -= text[j] == "}"
        if depth == 0:
            break
    os.makedirs(OUT, exist_ok=True)
    tu = "#include <cmath>\nnamespace Util {\n" + text[start:j + 1] + "\n}\n" \
         'extern "C" double ref_round(double v, unsigned char places) { return Util::round(v, places); }\n'
    with tempfile.TemporaryDirectory(prefix="hpo_ref_") as tmp:
        src = os.path.join(tmp, "unit_util.cpp")
        with open(src, "w") as f:
            f.write(tu)
        res = subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-w", "-shared", "-fPIC", src, "-o", UTIL_LIB],
                             capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("reference util shim build failed:\n%s" % res.stderr[-4000:])
    return UTIL_LIB


def build_all(variants=None, force=False):
    if not reference_available():
        return []
    variants = variants or DEFAULT_VARIANTS
    with ThreadPoolExecutor(max_workers=min(8, len(variants))) as ex:
        libs = list(ex.map(lambda v: build_variant(v, force), variants))
    return libs + [build_raster(force), build_util(force)]


if __name__ == "__main__":
    if not reference_available():
        print("reference tree not present at %s; nothing built" % REF_SRC)
        sys.exit(0)
    for path in build_all(sys.argv[1:] or None, force=True):
        print("built", os.path.relpath(path, os.path.dirname(HERE)))
