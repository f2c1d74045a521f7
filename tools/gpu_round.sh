#!/bin/bash
# One GPU visit for the record: parity tests, smoke, bench (f32 + f64 + reference arm), ncu launch list, ncu full captures.
# usage: tools/gpu_round.sh <tag>
tag=${1:-x}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/gpu.txt
python -m pytest tests -m gpu -q > $out/test.log 2>&1; echo "pytest rc=$?" >> $out/test.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1
python bench.py > $out/bench_f32.json 2> $out/bench_f32.err
python bench.py --dtype f64 > $out/bench_f64.json 2> $out/bench_f64.err
python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python bench.py --steps 3 --warmup 1 --no-cpu --e2e-steps 1 > $out/ncu_launch.log 2>&1
for k in vjp_march rhs_march; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o $out/prof_$k python bench.py --steps 3 --warmup 1 --no-cpu --e2e-steps 0 > $out/ncu_$k.log 2>&1
  ncu -i $out/prof_$k.ncu-rep --page raw --csv > $out/raw_$k.csv 2>/dev/null
  rm -f $out/prof_$k.ncu-rep
done
python tools/ncusum.py $out/raw_vjp_march.csv > $out/ncu_summary_f32.txt; python tools/ncusum.py $out/raw_rhs_march.csv >> $out/ncu_summary_f32.txt
tail -3 $out/test.log; cat $out/smoke.log; python tools/benchsum.py $out/bench_f32.json $out/bench_f64.json; cut -c1-260 $out/bench_ref.json
