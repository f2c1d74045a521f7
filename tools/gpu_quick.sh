#!/bin/bash
# quick GPU visit: parity tests + f32/f64 bench.   usage: tools/gpu_quick.sh <tag> [pytest -k expr]
tag=${1:-q}
out=gpurun_out/$tag
mkdir -p $out
if [ -n "$2" ]; then python -m pytest tests -m gpu -x -q -k "$2" > $out/test.log 2>&1; else python -m pytest tests -m gpu -x -q > $out/test.log 2>&1; fi
echo "pytest rc=$?" >> $out/test.log
tail -15 $out/test.log
python bench.py --cpu-seconds 3 > $out/bench_f32.json 2> $out/bench_f32.err
python bench.py --dtype f64 --no-cpu > $out/bench_f64.json 2> $out/bench_f64.err
python tools/benchsum.py $out/bench_f32.json $out/bench_f64.json
tail -3 $out/bench_f32.err
