#!/bin/bash
# ncu --set full of the fused RDPK stage launches (F1 + stage update: mode MID; reverse-ODE stage: interpolation + A1 + stage update).  usage: tools/gpu_ncu_rk.sh <tag>
out=gpurun_out/${1:-ncurk}; mkdir -p $out
B1='\(bool\)1'; B0='\(bool\)0'
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"sia2d_rhs_march2<$B1, $B0, $B1, $B0, \(int\)2>" -s 4 -c 1 -o $out/f1rk -f python tools/bench_rdpk.py f32 > $out/a.log 2>&1
ncu -i $out/f1rk.ncu-rep --page raw --csv > $out/raw_f1rk.csv 2>/dev/null; rm -f $out/f1rk.ncu-rep
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"sia2d_vjp_march2<$B1, $B0, $B1, $B0, $B1, $B0, $B0, $B1>" -s 12 -c 1 -o $out/rka -f python tools/bench_contadj.py f32 256 8 > $out/b.log 2>&1
ncu -i $out/rka.ncu-rep --page raw --csv > $out/raw_rka.csv 2>/dev/null; rm -f $out/rka.ncu-rep
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"sia2d_vjp_march<double, $B1, $B0, $B1, $B0, $B1, $B0, $B0, $B1>" -s 12 -c 1 -o $out/rka64 -f python tools/bench_contadj.py f64 256 8 > $out/c.log 2>&1
ncu -i $out/rka64.ncu-rep --page raw --csv > $out/raw_rka64.csv 2>/dev/null; rm -f $out/rka64.ncu-rep
python tools/ncu_summary.py $out/raw_f1rk.csv $out/raw_rka.csv $out/raw_rka64.csv > $out/summary.txt 2>&1
wc -c $out/*.csv; tail -3 $out/a.log
