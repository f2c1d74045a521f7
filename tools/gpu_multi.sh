#!/bin/bash
# Multi-GPU record on ONE box: bench.py (weak scaling, NCCL [loss; dtheta] all-reduce in the timed region, NCCL_DEBUG=INFO visible on stderr)
# and BASELINE configs 3 / 4 sharded over the ranks through the API mirror.   usage: tools/gpu_multi.sh <n_gpus> <outdir>
n=$1; out=${2:-gpurun_out/multi$n}
mkdir -p $out
export NCCL_DEBUG=INFO
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 "$@"; }
run bench.py --gpus $n --steps 20 --warmup 3 > $out/bench_${n}gpu.json 2> $out/bench_${n}gpu.err
run bench.py --impl reference --gpus $n --steps 20 --warmup 3 > $out/bench_reference_${n}gpu.json 2> $out/bench_reference_${n}gpu.err
run tools/bench_config3_mgpu.py > $out/configs_${n}gpu.jsonl 2> $out/configs_${n}gpu.err
wc -l $out/bench_${n}gpu.json; grep -c "nranks $n" $out/bench_${n}gpu.err; grep -h "^{" $out/configs_${n}gpu.jsonl | cut -c1-300
python tools/benchsum.py $out/bench_${n}gpu.json
