"""CPU-side checks of the drop-in boundary: the library builds for sm_100a, loads, exports every
symbol include/hipims_cuda.h declares, and refuses to run without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from hipims_ocl_b200 import build as hpbuild
from hipims_ocl_b200 import executor as hx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    hpbuild.build()
    return hx.load_library()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "hipims_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hp_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree(lib):
    declared = _declared_symbols()
    assert len(declared) >= 25
    assert sorted(hx.ABI) == declared
    for name in declared:
        assert getattr(lib, name) is not None


def test_abi_version_and_struct_sizes(lib):
    assert lib.hp_abi_version() == 2
    assert C.sizeof(hx.HpSchemeConfig) == 120
    assert C.sizeof(hx.HpSchemeStats) == 72
    assert C.sizeof(hx.HpBdyUniform) == 24 and C.sizeof(hx.HpBdyGridded) == 64 and C.sizeof(hx.HpBdyCell) == 40


def test_library_is_sm100a_only():
    out = os.popen("cuobjdump -lelf %s 2>/dev/null" % hx.LIB_PATH).read()
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_device_means_loud_failure(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(hx.HipimsCudaError, match="no CUDA device"):
        hx.device_count()
    with pytest.raises(hx.HipimsCudaError):
        hx.Executor(0)


def test_product_never_imports_the_oracle():
    """The oracle is a checker: nothing under hipims_ocl_b200/ or include/ may reference it."""
    bad = []
    for base in ("hipims_ocl_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            if "build" in dirpath.split(os.sep):
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                    text = open(os.path.join(dirpath, f), errors="ignore").read()
                    if re.search(r"(from|import)\s+oracle|oracle/|liboracle|hpo_", text):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad
