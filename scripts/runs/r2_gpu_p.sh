#!/bin/bash
# colsum folded into the weight-gradient GEMM, 6-stage ring for small single-pass GEMMs, plumbing audit, 64^3 latency floor
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_optim.py -q -x 2>&1 | tail -5 | cut -c1-250
timeout 300 python scripts/plumbing_audit.py > gpurun_out/r2p_audit.txt 2>&1; head -50 gpurun_out/r2p_audit.txt | cut -c1-220
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --no-attn-isolation --no-kernel-pass $EXTRA > gpurun_out/r2p_bench_$name.json 2> gpurun_out/r2p_bench_$name.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2p_bench_$name.json')); print('$name', d['ms_per_step'], d['gpu_launches'])
except Exception as e: print('$name failed', e)
PY
}
run base MICFORMER_FOLD_COLSUM=0 MICFORMER_GEMM_DEEP_RING=0
run fold MICFORMER_FOLD_COLSUM=1 MICFORMER_GEMM_DEEP_RING=0
run fold_ring MICFORMER_FOLD_COLSUM=1 MICFORMER_GEMM_DEEP_RING=1
EXTRA="--size 64" run size64 A=1
EXTRA="--size 32" run size32 A=1
