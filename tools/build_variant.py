#!/usr/bin/env python3
"""Development aid: build the fast-flavour kernels with extra -D flags into gpurun_out-independent
variant libraries (hipims_ocl_b200/variants/lib_<name>.so); select one with HIPIMS_CUDA_LIB=<path>.

    python tools/build_variant.py name -DHP_X=1 -DHP_Y=0
"""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hipims_ocl_b200 import build as B

name, defs = sys.argv[1], sys.argv[2:]
B.build()
out_dir = os.path.join(B.HERE, "variants")
os.makedirs(out_dir, exist_ok=True)
obj = os.path.join(out_dir, "hp_kernels_fast_%s.o" % name)
cmd = ["nvcc"] + B.ARCH + B.COMMON + ["-Xptxas", "-v", "-DHP_NS=hp_fast"] + defs + ["-c", os.path.join(B.CSRC, "hp_kernels.cu"), "-o", obj]
res = subprocess.run(cmd, capture_output=True, text=True)
if res.returncode:
    sys.exit(res.stderr[-4000:])
lines = res.stderr.splitlines()
for i, l in enumerate(lines):
    if "Compiling entry function" in l and ("march" in l) and "Id" in l:
        print(l.split("'")[1][:60], "|", lines[i + 2].strip(), "|", lines[i + 3].strip())
objs = [os.path.join(B.HERE, "build", f) for f in ("hp_kernels_strict.o", "hp_executor.o", "hp_comm.o")] + [obj]
lib = os.path.join(out_dir, "lib_%s.so" % name)
subprocess.check_call(["nvcc"] + B.ARCH + ["-shared", "-o", lib] + objs + ["-lcudart", "-ldl"])
print("built", lib)
