#!/bin/bash
# Build one tuning variant of the library (bench configuration only).  usage: tools/build_variant.sh <name> [-D...]
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$root/odinn.jl_b200/lib/var"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -DODINN_BENCH_ONLY "$@" \
    "$root"/odinn.jl_b200/csrc/*.cu -o "$root/odinn.jl_b200/lib/var/lib_$name.so"
