#!/bin/bash
# usage: tools/gpu_march.sh <tag> <march> [bench args]  -- parity tests + bench with ODINN_MARCH=<march>
tag=${1:-m}; export ODINN_MARCH=$2; shift; shift
out=gpurun_out/$tag
mkdir -p $out
timeout 300 python -m pytest tests -m gpu -x -q > $out/test.log 2>&1; echo "pytest rc=$?" >> $out/test.log
tail -12 $out/test.log
timeout 300 python bench.py --no-cpu --e2e-steps 0 "$@" > $out/bench_f32.json 2> $out/bench_f32.err
python tools/benchsum.py $out/bench_f32.json; tail -3 $out/bench_f32.err
