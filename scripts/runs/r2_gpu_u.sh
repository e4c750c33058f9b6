#!/bin/bash
# split-K forward GEMMs + ln_bwd without shared atomics: tests, bench A/B; ncu --set full of the tail's conv kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_parity2.py tests/test_gpu_optim.py tests/test_gpu_fused.py -q -x 2>&1 | tail -4 | cut -c1-250
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --no-attn-isolation --no-kernel-pass $EXTRA > gpurun_out/r2u_bench_$name.json 2> gpurun_out/r2u_bench_$name.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2u_bench_$name.json')); print('$name', round(d['ms_per_step'],3), d['gpu_launches'])
except Exception as e: print('$name failed', e)
PY
}
run nosplitk MICFORMER_FWD_SPLITK=0
run splitk MICFORMER_FWD_SPLITK=1
EXTRA="--size 64" run splitk_64 MICFORMER_FWD_SPLITK=1
python scripts/time_small.py 2>&1 | grep -v Warn | grep "ln_fwd"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv3_mma_bwd_weight|conv3_tc_kernel|block_permute" -c 6 -o gpurun_out/r2u_tail -f python scripts/prof_conv.py 1 > gpurun_out/r2u_ncu.log 2>&1; tail -3 gpurun_out/r2u_ncu.log
ls -la gpurun_out/r2u_tail.ncu-rep
