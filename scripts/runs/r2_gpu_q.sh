#!/bin/bash
# split-bf16 forward GEMMs (one-shot kernel) vs 3xTF32; colsum fold A/B repeated
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_parity2.py tests/test_gpu_optim.py -q -x 2>&1 | tail -5 | cut -c1-250
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --no-attn-isolation --no-kernel-pass $EXTRA > gpurun_out/r2q_bench_$name.json 2> gpurun_out/r2q_bench_$name.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2q_bench_$name.json')); print('$name', round(d['ms_per_step'],3), d['gpu_launches'])
except Exception as e: print('$name failed', e)
PY
}
run tf32x3_nofold MICFORMER_FWD_BF16X3=0 MICFORMER_FOLD_COLSUM=0
run tf32x3_fold MICFORMER_FWD_BF16X3=0 MICFORMER_FOLD_COLSUM=1
run bf3_nofold MICFORMER_FWD_BF16X3=1 MICFORMER_FOLD_COLSUM=0
run bf3_fold MICFORMER_FWD_BF16X3=1 MICFORMER_FOLD_COLSUM=1
run bf3_nofold_b MICFORMER_FWD_BF16X3=1 MICFORMER_FOLD_COLSUM=0
run bf3_fold_noring MICFORMER_FWD_BF16X3=1 MICFORMER_FOLD_COLSUM=1 MICFORMER_GEMM_DEEP_RING=0
EXTRA="--size 64" run bf3_size64 MICFORMER_FWD_BF16X3=1
