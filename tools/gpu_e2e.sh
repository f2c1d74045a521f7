#!/bin/bash
# e2e (host-batch) rate by pipeline chunk size.  usage: tools/gpu_e2e.sh <tag>
tag=${1:-e2e}; out=gpurun_out/$tag; mkdir -p $out
python -m pytest tests -m gpu -x -q -k "host_batch or mixed_size" 2>&1 | tail -2
for c in 2097152 4194304 8388608 16777216 33554432; do
  python bench.py --no-cpu --steps 10 --e2e-steps 8 --batch-chunk $c > $out/chunk$c.json 2> $out/chunk$c.err
  python tools/benchsum.py $out/chunk$c.json
done
