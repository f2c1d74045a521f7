#!/bin/bash
out=gpurun_out/${1:-cfg}; mkdir -p $out
python tools/bench_configs.py f32 > $out/configs_f32.jsonl 2> $out/configs_f32.err
python tools/bench_configs.py f64 > $out/configs_f64.jsonl 2> $out/configs_f64.err
cut -c1-330 $out/configs_f32.jsonl; cut -c1-330 $out/configs_f64.jsonl
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_rdpk.py -x -q -m gpu > $out/memcheck_rdpk.log 2>&1; echo "memcheck rc=$?" >> $out/memcheck_rdpk.log
tail -5 $out/memcheck_rdpk.log
