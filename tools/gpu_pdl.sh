#!/bin/bash
# PDL (programmatic dependent launch) of the F1 kernels: parity tests, A/B on the mid-size ensembles and the bench line.  usage: tools/gpu_pdl.sh <tag>
out=gpurun_out/${1:-pdl}
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/test.log 2>&1; echo "pytest rc=$?" >> $out/test.log
tail -4 $out/test.log
for v in 0 1; do
  echo "== ODINN_PDL=$v" | tee -a $out/pdl_ab.txt
  ODINN_PDL=$v python tools/bench_config4_fwd.py f32 32 | tee -a $out/pdl_ab.txt
  ODINN_PDL=$v python tools/bench_config4_fwd.py f32 64 | tee -a $out/pdl_ab.txt
  ODINN_PDL=$v python tools/bench_config4_fwd.py f64 64 | tee -a $out/pdl_ab.txt
  ODINN_PDL=$v python tools/bench_timeloop.py f32 2>&1 | grep -i "forward" | tee -a $out/pdl_ab.txt
  ODINN_PDL=$v python tools/bench_rdpk.py f32 | tee -a $out/pdl_ab.txt
done
python bench.py --dtype f64 --no-cpu --no-grad --no-other-dtype --e2e-steps 0 > $out/bench_f64.json 2> $out/bench_f64.err
python tools/benchsum.py $out/bench_f64.json
