#!/bin/bash
# ncu --set full capture of ONE launch of the fused step kernel: marching (ODINN_MARCH=2) and 2-D TMA ring (ODINN_MARCH=4) variants.
# usage: tools/ncu_fused.sh <outdir>
out=${1:-gpurun_out/ncu}
mkdir -p $out
run() {  # $1 march, $2 kernel regex, $3 launches to skip
  ODINN_MARCH=$1 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c 1 \
      -o $out/fused_m$1 -f python bench.py --steps 4 --warmup 2 --no-cpu --no-grad --no-other-dtype --e2e-steps 0 > $out/ncu_m$1.log 2>&1
  ncu -i $out/fused_m$1.ncu-rep --page raw --csv > $out/raw_m$1.csv 2>/dev/null
  rm -f $out/fused_m$1.ncu-rep   # (gpurun brings back at most 64 MiB: the summaries are what is kept)
}
run 2 sia2d_vjp_march2 2   # launches 0, 2, 4, ... of the marching kernel are fused steps (1, 3: the S-only pass of the iteration boundary)
run 4 sia2d_fused_tma 1
