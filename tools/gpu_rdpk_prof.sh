#!/bin/bash
# ncu launch list of the RDPK3Sp35 engine at the bench workload + one --set full capture of a fused stage launch.  usage: tools/gpu_rdpk_prof.sh <tag>
out=gpurun_out/${1:-rkprof}
mkdir -p $out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $out/launches_fused.csv python tools/bench_rdpk.py f32 > $out/l1.log 2>&1
ODINN_RK_NO_FUSE=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $out/launches_unfused.csv python tools/bench_rdpk.py f32 > $out/l2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sia2d_rhs_march2 -s 12 -c 2 -o $out/stage -f python tools/bench_rdpk.py f32 > $out/l3.log 2>&1
ncu -i $out/stage.ncu-rep --page raw --csv > $out/raw_stage.csv 2>/dev/null
ncu -i $out/stage.ncu-rep --page source --csv > $out/src_stage.csv 2>/dev/null
python tools/ncusum.py $out/raw_stage.csv > $out/stage_summary.txt
rm -f $out/stage.ncu-rep
