#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -x -m gpu 2>&1 | tail -3 | cut -c1-250
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --no-attn-isolation --no-kernel-pass $EXTRA > gpurun_out/r2z_bench_$name.json 2> gpurun_out/r2z_bench_$name.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2z_bench_$name.json')); print('$name', round(d['ms_per_step'],3), d['gpu_launches'])
except Exception as e: print('$name failed', e)
PY
}
run cat0 MICFORMER_QKV_CAT=0
run cat1 MICFORMER_QKV_CAT=1
run cat0b MICFORMER_QKV_CAT=0
run cat1b MICFORMER_QKV_CAT=1
