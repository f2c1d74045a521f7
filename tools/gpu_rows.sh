#!/bin/bash
tag=${1:-rows}; lib=$2; shift; shift
out=gpurun_out/$tag
mkdir -p $out
for r in 8 16 32 64 128; do
  ODINN_CHUNK_ROWS2=$r ODINN_B200_LIB=$PWD/$lib python bench.py --no-cpu --e2e-steps 0 --steps 30 "$@" > $out/rows$r.json 2> $out/rows$r.err
  python tools/benchsum.py $out/rows$r.json
done
