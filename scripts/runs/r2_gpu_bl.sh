#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_parity2.py tests/test_gpu_callers.py -q -x -k "whole_model or full_size or known_answers or tiny64 or batch or loss_curve or linear or epilogues or state_dict or callers or train" 2>&1 | tail -4 | cut -c1-250
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --no-attn-isolation --no-kernel-pass > gpurun_out/r2bl_bench_$name.json 2> gpurun_out/r2bl_bench_$name.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2bl_bench_$name.json')); print('$name', round(d['ms_per_step'],3), d['gpu_launches'])
except Exception as e: print('$name failed', e); print(open('gpurun_out/r2bl_bench_$name.err').read()[-1500:])
PY
}
run view MICFORMER_TAIL_VIEW=1
run noview MICFORMER_TAIL_VIEW=0
run view2 MICFORMER_TAIL_VIEW=1
