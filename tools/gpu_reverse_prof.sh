#!/bin/bash
# per-kernel durations of the reverse loop (fused and three-pass forms) and of an adaptive trial step at the bench workload
out=${1:-gpurun_out/rev}
mkdir -p $out
python tools/bench_timeloop.py f32 > $out/timeloop_f32.jsonl 2>&1
ODINN_NO_FUSE=1 python tools/bench_timeloop.py f32 > $out/timeloop_f32_nofuse.jsonl 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_timeloop.csv python tools/bench_timeloop.py f32 > $out/ncu_timeloop.log 2>&1
ODINN_NO_FUSE=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_timeloop_nofuse.csv python tools/bench_timeloop.py f32 > $out/ncu_timeloop_nofuse.log 2>&1
python tools/bench_adaptive.py f32 > $out/adaptive_f32.jsonl 2>&1
cat $out/timeloop_f32.jsonl $out/timeloop_f32_nofuse.jsonl $out/adaptive_f32.jsonl | cut -c1-300
