"""Builds libhipims_cuda.so (the C-ABI CUDA executor) in-tree for sm_100a.

    python -m hipims_ocl_b200.build [--force]

nvcc cross-compiles without a GPU.  hp_kernels.cu is compiled twice: strict (-fmad=false) and
fast (FMA contraction); see csrc/hp_kernels.cu.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB = os.path.join(HERE, "libhipims_cuda.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", INCLUDE, "-I", CSRC]


def _nccl_include():
    for cand in ("/usr/include",):
        if os.path.exists(os.path.join(cand, "nccl.h")):
            return cand
    try:
        import nvidia.nccl
        return os.path.join(os.path.dirname(nvidia.nccl.__file__), "include")
    except Exception:
        return None


def _sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(INCLUDE, "hipims_cuda.h"),
                                                                        os.path.abspath(__file__)]


def up_to_date():
    return os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in _sources())


def _run(cmd):
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("build step failed:\n%s\n%s" % (" ".join(cmd), (res.stdout + res.stderr)[-6000:]))
    return res.stdout + res.stderr


def build(force=False, verbose=False):
    if not force and up_to_date():
        return LIB
    obj = os.path.join(HERE, "build")
    os.makedirs(obj, exist_ok=True)
    ptxas = ["-Xptxas", "-v"] if verbose else []
    nccl_inc = _nccl_include()
    jobs = [
        ["nvcc"] + ARCH + COMMON + ptxas + ["-DHP_NS=hp_strict", "-DHP_FLAVOUR_STRICT", "-fmad=false", "-c",
                                            os.path.join(CSRC, "hp_kernels.cu"), "-o", os.path.join(obj, "hp_kernels_strict.o")],
        ["nvcc"] + ARCH + COMMON + ptxas + ["-DHP_NS=hp_fast"] + (["-DHP_GODUNOV_STAGES=" + os.environ["HP_GODUNOV_STAGES"]] if "HP_GODUNOV_STAGES" in os.environ else []) + ["-c", os.path.join(CSRC, "hp_kernels.cu"), "-o",
                                            os.path.join(obj, "hp_kernels_fast.o")],
        ["nvcc"] + ARCH + COMMON + ["-c", os.path.join(CSRC, "hp_executor.cu"), "-o", os.path.join(obj, "hp_executor.o")],
        ["nvcc"] + ARCH + COMMON + (["-I", nccl_inc] if nccl_inc else []) + ["-x", "cu", "-c", os.path.join(CSRC, "hp_comm.cpp"),
                                                                              "-o", os.path.join(obj, "hp_comm.o")],
    ]
    with ThreadPoolExecutor(max_workers=4) as ex:
        logs = list(ex.map(_run, jobs))
    if verbose:
        print("\n".join(logs))
    objs = [os.path.join(obj, f) for f in ("hp_kernels_strict.o", "hp_kernels_fast.o", "hp_executor.o", "hp_comm.o")]
    _run(["nvcc"] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart", "-ldl"])
    return LIB


HOST = os.path.join(HERE, "host")
HOST_LIB = os.path.join(HERE, "libhipims_host.so")
HOST_EXE = os.path.join(HERE, "hipims-b200")


def build_host(force=False):
    """C++ host mirror (CScheme*, CDomainCartesian, CBoundary*, XML) above the C ABI + the CLI."""
    srcs = [os.path.join(HOST, f) for f in ("hipims_host.cpp", "hipims_host.h", "main.cpp")] + [os.path.join(INCLUDE, "hipims_cuda.h")]
    if not force and os.path.exists(HOST_LIB) and os.path.exists(HOST_EXE) and \
            all(min(os.path.getmtime(HOST_LIB), os.path.getmtime(HOST_EXE)) >= os.path.getmtime(s) for s in srcs + [LIB]):
        return HOST_LIB
    common = ["g++", "-std=c++17", "-O2", "-Wall", "-fPIC", "-I", INCLUDE]
    link = ["-L", HERE, "-lhipims_cuda", "-Wl,-rpath,$ORIGIN"]
    _run(common + ["-shared", os.path.join(HOST, "hipims_host.cpp"), "-o", HOST_LIB] + link)
    _run(common + [os.path.join(HOST, "main.cpp"), os.path.join(HOST, "hipims_host.cpp"), "-o", HOST_EXE] + link)
    return HOST_LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built", path)
    print("built", build_host(force="--force" in sys.argv))
