#!/bin/bash
out=gpurun_out/${1:-ca}; mkdir -p $out
python -m pytest tests -m gpu -x -q -k "rdpk or continuous or adjoint or next" > $out/test.log 2>&1; echo "pytest rc=$?" >> $out/test.log; tail -5 $out/test.log
for d in f32 f64; do
  python tools/bench_contadj.py $d | tee -a $out/ca.jsonl
  ODINN_RK_NO_FUSE=1 python tools/bench_contadj.py $d | tee -a $out/ca.jsonl
done
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file $out/launches_f32.csv python tools/bench_contadj.py f32 64 8 > $out/l1.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file $out/launches_f64.csv python tools/bench_contadj.py f64 64 8 > $out/l2.log 2>&1
