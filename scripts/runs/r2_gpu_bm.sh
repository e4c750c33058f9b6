#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --no-attn-isolation --no-kernel-pass "$@" > gpurun_out/r2bm_bench_$name.json 2> gpurun_out/r2bm_bench_$name.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2bm_bench_$name.json')); print('$name', round(d['ms_per_step'],3), d['gpu_launches'])
except Exception as e: print('$name failed', e); print(open('gpurun_out/r2bm_bench_$name.err').read()[-1500:])
PY
}
run arena0 --arena 0
run arena1 --arena 1
run arena0b --arena 0
run arena1b --arena 1
