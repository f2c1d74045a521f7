#!/bin/bash
out=gpurun_out/${1:-final}; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/test.log 2>&1; echo "pytest rc=$?" >> $out/test.log; tail -5 $out/test.log
python bench.py > $out/bench_f32.json 2> $out/bench_f32.err; python tools/benchsum.py $out/bench_f32.json
python bench.py --dtype f64 --no-cpu --no-grad --no-other-dtype > $out/bench_f64.json 2> $out/bench_f64.err; python tools/benchsum.py $out/bench_f64.json
for d in f32 f64; do python tools/bench_timeloop.py $d 2>/dev/null | grep -E "forward|reverse" | cut -c1-230 | tee -a $out/timeloop.jsonl; python tools/bench_rdpk.py $d | cut -c1-200 | tee -a $out/timeloop.jsonl; python tools/bench_contadj.py $d | cut -c1-330 | tee -a $out/timeloop.jsonl; done
