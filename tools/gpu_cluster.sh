#!/bin/bash
# cluster-resident forward solves: parity tests + configs 1 / 3 / 4 with the cluster path off / automatic / forced sizes.   usage: tools/gpu_cluster.sh <tag> [modes]
tag=${1:-cl}
modes=${2:-"0 -1"}
out=gpurun_out/$tag
mkdir -p $out
timeout 1200 python -m pytest tests/test_gpu_cluster.py tests/test_gpu_rdpk.py tests/test_gpu_timeloop.py tests/test_gpu_next.py -x -q > $out/test.log 2>&1
echo "pytest rc=$?" >> $out/test.log
tail -25 $out/test.log
for m in $modes; do
  for dt in f32 f64; do
    ODINN_CLUSTER=$m timeout 600 python tools/bench_configs.py $dt > $out/configs_${dt}_cl$m.jsonl 2> $out/configs_${dt}_cl$m.err
    echo "== ODINN_CLUSTER=$m $dt"; cut -c1-300 $out/configs_${dt}_cl$m.jsonl; tail -2 $out/configs_${dt}_cl$m.err
  done
done
