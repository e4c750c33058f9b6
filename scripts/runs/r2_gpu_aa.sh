#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity2.py tests/test_gpu_multi.py tests/test_gpu_optim.py tests/test_gpu_callers.py -q -x 2>&1 | tail -3 | cut -c1-250
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --no-attn-isolation --no-kernel-pass $EXTRA > gpurun_out/r2aa_bench_$name.json 2> gpurun_out/r2aa_bench_$name.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2aa_bench_$name.json')); print('$name', round(d['ms_per_step'],3), d['gpu_launches'])
except Exception as e: print('$name failed', e)
PY
}
run q0 MICFORMER_Q_SIDE=0
run q1 MICFORMER_Q_SIDE=1
run q0b MICFORMER_Q_SIDE=0
run q1b MICFORMER_Q_SIDE=1
