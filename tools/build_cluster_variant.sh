#!/bin/bash
# Rebuild launch_cluster.cu with extra -D flags and link it with the objects of the regular build.  usage: tools/build_cluster_variant.sh <name> [-D...]
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$root/odinn.jl_b200/lib/var"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c "$root/odinn.jl_b200/csrc/launch_cluster.cu" -o "/tmp/launch_cluster_$name.o" || exit 1
objs=$(ls "$root"/odinn.jl_b200/lib/obj/*.o | grep -v launch_cluster.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC $objs "/tmp/launch_cluster_$name.o" -o "$root/odinn.jl_b200/lib/var/lib_$name.so" -lcuda
