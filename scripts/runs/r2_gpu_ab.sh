#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -q -x -k "conv3 or golden or full_size or tc_whole or seg or tail" 2>&1 | tail -3 | cut -c1-250
echo "row-owner dW8:"; python scripts/prof_conv.py 20 2>&1 | grep out_conv
echo "old dW8:"; MICFORMER_CONV_DW8=0 python scripts/prof_conv.py 20 2>&1 | grep out_conv
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --no-attn-isolation --no-kernel-pass $EXTRA > gpurun_out/r2ab_bench_$name.json 2> gpurun_out/r2ab_bench_$name.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2ab_bench_$name.json')); print('$name', round(d['ms_per_step'],3), d['gpu_launches'])
except Exception as e: print('$name failed', e)
PY
}
run dw8_0 MICFORMER_CONV_DW8=0
run dw8_1 MICFORMER_CONV_DW8=1
