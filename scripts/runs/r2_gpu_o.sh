#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -q -x -k "conv3 or offset or golden or tc_whole or full_size" 2>&1 | tail -5 | cut -c1-250
for v in 0 1; do
MICFORMER_CONV_TMA=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --no-attn-isolation --dump-kernels gpurun_out/r2o_kernels_$v.json > gpurun_out/r2o_bench_$v.json 2> gpurun_out/r2o_bench_$v.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2o_bench_$v.json')); print('conv_tma=$v', d['ms_per_step'])
k=json.load(open('gpurun_out/r2o_kernels_$v.json'))
for r in k['kernels']:
    if 'conv3_tc' in r['kernel']: print(f"   {r['kernel']:52s} {r['ms_per_step']:.3f} {r['calls_per_step']:5.0f} {r['us_per_call']:7.1f}")
PY
done
