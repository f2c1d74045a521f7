#!/bin/bash
# Time the fused step for the marching kernel and the TMA ring variants (ODINN_TMA_VARIANT: 0 = 4 stages / 5 CTAs, 1 = 3 / 6, 2 = 3 / 5, 3 = 6 / 5).
out=${1:-gpurun_out/sweep_tma}
mkdir -p $out
B="python bench.py --steps 40 --warmup 5 --no-cpu --no-grad --no-other-dtype --e2e-steps 0"
ODINN_MARCH=2 $B > $out/m2.json 2> $out/m2.err
for v in 0 1 2 3; do ODINN_MARCH=4 ODINN_TMA_VARIANT=$v $B > $out/tma_v$v.json 2> $out/tma_v$v.err; done
python - <<PY
import json, glob
for f in sorted(glob.glob("$out/*.json")):
    try:
        d = json.load(open(f)); print(f, "%.4f ms  frac %.4f" % (d["roofline"]["ms_per_launch"], d["roofline"]["frac"]))
    except Exception as ex:
        print(f, "failed", ex)
PY
