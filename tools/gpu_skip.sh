#!/bin/bash
out=gpurun_out/${1:-skip}; mkdir -p $out
python -m pytest tests -m gpu -x -q -k "rdpk or continuous or adjoint or next or config" > $out/test.log 2>&1; echo "pytest rc=$?" >> $out/test.log; tail -4 $out/test.log
for d in f32 f64; do
  python tools/bench_rdpk.py $d | tee -a $out/rdpk.jsonl
  python tools/bench_contadj.py $d | tee -a $out/ca.jsonl
done
