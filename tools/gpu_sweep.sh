#!/bin/bash
# bench every library variant under odinn.jl_b200/lib/var (ODINN_B200_LIB override).  usage: tools/gpu_sweep.sh <tag> [bench args]
tag=${1:-sw}; shift
out=gpurun_out/$tag
mkdir -p $out
for f in odinn.jl_b200/lib/var/*.so; do
  n=$(basename $f .so)
  ODINN_B200_LIB=$PWD/$f python bench.py --no-cpu --e2e-steps 0 --steps 30 "$@" > $out/$n.json 2> $out/$n.err
  python tools/benchsum.py $out/$n.json
done
