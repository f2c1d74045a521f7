#!/bin/bash
# Rebuild ONE translation unit of csrc/ with extra -D flags and link it with the objects of the regular build.
# usage: tools/build_tu_variant.sh <tu without .cu> <name> [-D...]   ->  odinn.jl_b200/lib/var/lib_<name>.so  (use with ODINN_B200_LIB=...)
tu=$1; name=$2; shift 2
root=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$root/odinn.jl_b200/lib/var"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c "$root/odinn.jl_b200/csrc/$tu.cu" -o "/tmp/${tu}_$name.o" || exit 1
objs=$(ls "$root"/odinn.jl_b200/lib/obj/*.o | grep -v "/$tu.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC $objs "/tmp/${tu}_$name.o" -o "$root/odinn.jl_b200/lib/var/lib_$name.so" -lcuda
