#!/bin/bash
# fp64 bench of one library variant, fused and unfused.  usage: tools/gpu_var64.sh <tag> <lib.so>
tag=${1:-v}; lib=$2; out=gpurun_out/$tag; mkdir -p $out
ODINN_B200_LIB=$PWD/$lib python bench.py --dtype f64 --no-cpu --e2e-steps 0 --steps 30 > $out/fused.json 2> $out/fused.err
ODINN_B200_LIB=$PWD/$lib python bench.py --dtype f64 --no-cpu --e2e-steps 0 --steps 30 --no-fuse > $out/unfused.json 2> $out/unfused.err
python tools/benchsum.py $out/fused.json $out/unfused.json; tail -3 $out/fused.err
