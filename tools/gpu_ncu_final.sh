#!/bin/bash
# ncu evidence of the final build: launch list of the bench command, --set full of the fused step (fp32 TMA ring with 126-row bands,
# fp64 with 100-row chunks) -> <out>/launches.csv, <out>/ncu_summary.txt.   usage: tools/gpu_ncu_final.sh <tag>
out=gpurun_out/${1:-rec7}; mkdir -p $out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-other-dtype > $out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sia2d_fused_tma -s 1 -c 1 -o $out/tma -f python bench.py --steps 4 --warmup 2 --no-cpu --no-grad --no-other-dtype --e2e-steps 0 > $out/a.log 2>&1
ncu -i $out/tma.ncu-rep --page raw --csv > $out/raw_tma.csv 2>/dev/null; rm -f $out/tma.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:sia2d_vjp_march -s 2 -c 1 -o $out/f64 -f python bench.py --dtype f64 --steps 4 --warmup 2 --no-cpu --no-grad --no-other-dtype --e2e-steps 0 > $out/b.log 2>&1
ncu -i $out/f64.ncu-rep --page raw --csv > $out/raw_f64.csv 2>/dev/null; rm -f $out/f64.ncu-rep
python tools/ncu_summary.py $out/raw_tma.csv $out/raw_f64.csv > $out/ncu_summary.txt 2>&1
grep -E "^----|gpu__time_duration|dram__bytes" $out/ncu_summary.txt
