#!/bin/bash
# ncu --set full capture of ONE launch of a kernel inside the bench step.  usage: tools/ncu_one.sh <outdir> <tag> <kernel regex> <skip> [env assignments...]
out=$1; tag=$2; rx=$3; skip=$4; shift 4
mkdir -p $out
env "$@" ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c 1 -o $out/$tag -f \
    python bench.py --steps 4 --warmup 2 --no-cpu --no-grad --no-other-dtype --e2e-steps 0 > $out/ncu_$tag.log 2>&1
ncu -i $out/$tag.ncu-rep --page raw --csv > $out/raw_$tag.csv 2>/dev/null
ncu -i $out/$tag.ncu-rep --page source --csv > $out/src_$tag.csv 2>/dev/null
