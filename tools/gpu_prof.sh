#!/bin/bash
# ncu full capture of one launch of each marching kernel inside a short bench run; raw CSV pages are made on the box
# (the .ncu-rep of the A1+A2 / fused kernel is kept, gpurun_out/ is capped at 64 MiB).
# usage: tools/gpu_prof.sh <tag> [extra bench args]
tag=${1:-p}; shift
out=gpurun_out/$tag
mkdir -p $out
for k in vjp_march rhs_march; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o $out/prof_$k python bench.py --steps 3 --warmup 1 --no-cpu --e2e-steps 0 "$@" > $out/ncu_$k.log 2>&1
  ncu -i $out/prof_$k.ncu-rep --page raw --csv > $out/raw_$k.csv 2>/dev/null
done
rm -f $out/prof_rhs_march.ncu-rep
python tools/ncusum.py $out/raw_vjp_march.csv > $out/summary.txt; python tools/ncusum.py $out/raw_rhs_march.csv >> $out/summary.txt
cat $out/summary.txt | cut -c1-150
