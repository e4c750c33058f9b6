#!/bin/bash
mkdir -p gpurun_out
for v in 0 192 384 64; do
MICFORMER_FUSED_SPLIT=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --no-attn-isolation --no-kernel-pass > gpurun_out/r2n_bench_$v.json 2> gpurun_out/r2n_bench_$v.err
python -c "
import json; d=json.load(open('gpurun_out/r2n_bench_$v.json')); print('split>=$v', d['ms_per_step'], d['gpu_launches']/d['steps'])"
done
