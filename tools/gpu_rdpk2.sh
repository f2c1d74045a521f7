#!/bin/bash
# RDPK3Sp35 fused-stage engine: parity tests, A/B timing, launch list.  usage: tools/gpu_rdpk2.sh <tag>
out=gpurun_out/${1:-rdpk}
mkdir -p $out
python -m pytest tests -m gpu -x -q -k "rdpk or adaptive or config1 or config3" > $out/test.log 2>&1; echo "pytest rc=$?" >> $out/test.log
tail -4 $out/test.log
for d in f32 f64; do
  python tools/bench_rdpk.py $d | tee -a $out/rdpk.jsonl
  ODINN_RK_NO_FUSE=1 python tools/bench_rdpk.py $d | tee -a $out/rdpk.jsonl
done
for d in f32 f64; do
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $out/launches_fused_$d.csv python tools/bench_rdpk.py $d > $out/l1.log 2>&1
done
