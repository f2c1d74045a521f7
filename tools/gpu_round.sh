#!/bin/bash
# One GPU visit: parity tests, bench (f32 + f64), ncu launch list, ncu full capture of the two marching kernels.
# usage: tools/gpu_round.sh <tag>
tag=${1:-x}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/gpu.txt
python -m pytest tests -m gpu -x -q > $out/test.log 2>&1; echo "pytest rc=$?" >> $out/test.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1
python bench.py > $out/bench_f32.json 2> $out/bench_f32.err
python bench.py --dtype f64 > $out/bench_f64.json 2> $out/bench_f64.err
python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python bench.py --steps 3 --warmup 1 --no-cpu --e2e-steps 1 > $out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:march -s 2 -c 2 -o $out/prof_full python bench.py --steps 3 --warmup 1 --no-cpu --e2e-steps 0 > $out/ncu_full.log 2>&1
tail -3 $out/test.log; cat $out/smoke.log; cat $out/bench_f32.json
