#!/bin/bash
out=gpurun_out/${1:-sg}; mkdir -p $out
python -m pytest tests -m gpu -x -q -k "f64 or parity or baseline" > $out/test.log 2>&1; echo "pytest rc=$?" >> $out/test.log; tail -3 $out/test.log
for r in 1 2; do
python bench.py --dtype f64 --no-cpu --no-grad --no-other-dtype --e2e-steps 0 --steps 30 > $out/xy1_$r.json 2> $out/xy1.err
ODINN_B200_LIB=$PWD/odinn.jl_b200/lib/var/lib_sg0.so python bench.py --dtype f64 --no-cpu --no-grad --no-other-dtype --e2e-steps 0 --steps 30 > $out/xy0_$r.json 2> $out/xy0.err
done
python tools/benchsum.py $out/xy1_1.json $out/xy0_1.json $out/xy1_2.json $out/xy0_2.json
