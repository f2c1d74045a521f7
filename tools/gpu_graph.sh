#!/bin/bash
out=gpurun_out/${1:-graph}; mkdir -p $out
python -m pytest tests -m gpu -x -q -k "rdpk or config or api" > $out/test.log 2>&1; echo "pytest rc=$?" >> $out/test.log; tail -4 $out/test.log
for v in 1 0; do
  echo "== ODINN_RK_GRAPH=$v" | tee -a $out/ab.txt
  ODINN_RK_GRAPH=$v python tools/bench_configs.py f32 2>/dev/null | grep -E "RDPK|DEFAULT" | grep -E '"(3|4):' | cut -c1-260 | tee -a $out/ab.txt
  ODINN_RK_GRAPH=$v python tools/bench_configs.py f64 2>/dev/null | grep -E "RDPK|DEFAULT" | grep -E '"(3|4):' | cut -c1-260 | tee -a $out/ab.txt
done
