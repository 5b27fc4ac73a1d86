#!/bin/bash
# tools/ab.sh <workloads> <steps> <lib> [<lib> ...]  -- development aid: tools/quick_bench.py under several builds of the library
wl=$1; steps=$2; shift 2
for lib in "$@"; do
    echo "== $lib"
    if [ "$lib" = main ]; then python tools/quick_bench.py "$wl" "$steps"; else HIPIMS_CUDA_LIB=hipims_ocl_b200/variants/lib_$lib.so python tools/quick_bench.py "$wl" "$steps"; fi
done
