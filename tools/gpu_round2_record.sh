#!/bin/bash
# The measured record of a build: GPU tests, bench lines (both arms, f32 + f64), ncu launch list of the bench command, ncu --set full
# of the dominant kernel (marching and TMA variants).  usage: tools/gpu_round2_record.sh <outdir>
out=${1:-gpurun_out/record}
mkdir -p $out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem --format=csv > $out/gpu.txt; nproc >> $out/gpu.txt
(time python -m pytest tests -m gpu -q --durations=10) > $out/gpu_tests.log 2>&1
python bench.py --impl reference > $out/bench_reference.json 2> $out/bench_reference.err
python bench.py > $out/bench_f32.json 2> $out/bench_f32.err
ODINN_MARCH=2 python bench.py --no-cpu --no-grad --no-other-dtype --e2e-steps 0 > $out/bench_f32_march2.json 2> $out/bench_f32_march2.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-other-dtype > $out/launches_bench.log 2>&1
tools/ncu_fused.sh $out > /dev/null 2>&1
python tools/ncu_summary.py $out/raw_m2.csv $out/raw_m4.csv > $out/ncu_summary_fused.txt
