#!/bin/bash
# RDPK3Sp35 fused-stage engine: parity tests + A/B timing.  usage: tools/gpu_rdpk.sh <tag>
out=gpurun_out/${1:-rdpk}
mkdir -p $out
python -m pytest tests -m gpu -x -q -k "rdpk or adaptive or config1 or config3" > $out/test.log 2>&1; echo "pytest rc=$?" >> $out/test.log
tail -8 $out/test.log
for d in f32 f64; do
  python tools/bench_rdpk.py $d | tee -a $out/rdpk.jsonl
  ODINN_RK_NO_FUSE=1 python tools/bench_rdpk.py $d | tee -a $out/rdpk.jsonl
done
