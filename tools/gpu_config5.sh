#!/bin/bash
# BASELINE config 5: HBM roofline sweep of the stencil-only forward (F1) on 2048x2048 grids, fp32 and fp64:
# bench numbers (CUDA events) for 1 grid (L2-resident: 16 / 32 MB per plane) and 32 grids (> L2), plus one ncu pass with DRAM bytes.
tag=${1:-c5}; out=gpurun_out/$tag; mkdir -p $out
for dt in f32 f64; do
  for G in 1 32; do
    python bench.py --grid 2048 --glaciers $G --dtype $dt --no-cpu --e2e-steps 0 --steps 30 > $out/bench_${dt}_G$G.json 2> $out/bench_${dt}_G$G.err
    python tools/benchsum.py $out/bench_${dt}_G$G.json
  done
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:rhs_march -s 2 -c 3 --csv \
      --log-file $out/ncu_rhs_$dt.csv python bench.py --grid 2048 --glaciers 32 --dtype $dt --no-cpu --e2e-steps 0 --steps 3 --warmup 1 > /dev/null 2>&1
  grep -v "^==" $out/ncu_rhs_$dt.csv | cut -d, -f5,10- | tail -9
done
