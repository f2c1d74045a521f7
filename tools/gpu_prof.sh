#!/bin/bash
# ncu full capture of the marching kernels inside a short bench run.  usage: tools/gpu_prof.sh <tag> [extra bench args]
tag=${1:-p}; shift
out=gpurun_out/$tag
mkdir -p $out
ncu --set full --clock-control none --import-source on -k regex:march -s 2 -c 2 -o $out/prof_full python bench.py --steps 3 --warmup 1 --no-cpu --e2e-steps 0 "$@" > $out/ncu_full.log 2>&1
tail -2 $out/ncu_full.log | cut -c1-300
