#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -q -x -k "conv3 or offset or golden or block or full_size" 2>&1 | tail -3 | cut -c1-250
for pct in 200 100 150 300; do echo "target pct $pct"; MICFORMER_CONV_TARGET_PCT=$pct python scripts/prof_conv.py 20 2>&1 | grep -v Warn; done
python scripts/time_small.py 2>&1 | grep -v Warn
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --no-attn-isolation --no-kernel-pass $EXTRA > gpurun_out/r2v_bench_$name.json 2> gpurun_out/r2v_bench_$name.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2v_bench_$name.json')); print('$name', round(d['ms_per_step'],3), d['gpu_launches'])
except Exception as e: print('$name failed', e)
PY
}
run t200 MICFORMER_CONV_TARGET_PCT=200
run t100 MICFORMER_CONV_TARGET_PCT=100
