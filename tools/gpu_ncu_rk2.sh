#!/bin/bash
out=gpurun_out/${1:-ncurk}; mkdir -p $out
B1='\(bool\)1'; B0='\(bool\)0'
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"sia2d_vjp_march2<$B1, $B0, $B1, $B0, $B1, $B0, $B0, $B1>" -s 1 -c 1 -o $out/rka -f python tools/bench_contadj.py f32 256 8 > $out/b.log 2>&1
ncu -i $out/rka.ncu-rep --page raw --csv > $out/raw_rka.csv 2>/dev/null; rm -f $out/rka.ncu-rep
python tools/ncu_summary.py $out/raw_f1rk.csv $out/raw_rka.csv $out/raw_rka64.csv > $out/summary.txt 2>&1
grep -E "^----|gpu__time_duration|dram__bytes" $out/summary.txt
