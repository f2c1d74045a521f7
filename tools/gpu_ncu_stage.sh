#!/bin/bash
# ncu --set full of one SSPRK3 stage launch (F1 + fused stage, long-chunk table) and one plain F1 launch at the bench workload.  usage: tools/gpu_ncu_stage.sh <tag>
out=gpurun_out/${1:-ncustage}; mkdir -p $out
B1='\(bool\)1'; B0='\(bool\)0'
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"sia2d_rhs_march2<$B1, $B0, $B1, $B1, \(int\)0>" -s 30 -c 1 -o $out/stage -f python tools/bench_timeloop.py f32 > $out/a.log 2>&1
ncu -i $out/stage.ncu-rep --page raw --csv > $out/raw_stage.csv 2>/dev/null; rm -f $out/stage.ncu-rep
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"sia2d_rhs_march2<$B1, $B0, $B1, $B0, \(int\)0>" -s 3 -c 1 -o $out/f1 -f python bench.py --steps 4 --warmup 2 --no-cpu --no-grad --no-other-dtype --e2e-steps 0 > $out/b.log 2>&1
ncu -i $out/f1.ncu-rep --page raw --csv > $out/raw_f1.csv 2>/dev/null; rm -f $out/f1.ncu-rep
python tools/ncu_summary.py $out/raw_stage.csv $out/raw_f1.csv > $out/summary.txt 2>&1
grep -E "^----|gpu__time_duration|dram__bytes|launch__grid_size|issue_active|long_scoreboard" $out/summary.txt
