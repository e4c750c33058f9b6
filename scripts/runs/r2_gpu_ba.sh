#!/bin/bash
# first run of the tcgen05 window-attention backward: numerics vs fp64 / CUDA-core kernel, config-4 timing, tests, w7 step
mkdir -p gpurun_out
timeout 240 python scripts/check_tc_attn_bwd.py > gpurun_out/r2ba_check.log 2>&1; echo "check rc=$?"; tail -12 gpurun_out/r2ba_check.log | cut -c1-220
timeout 400 python -m pytest tests/test_gpu_tc.py -q -x -k "window_attention or w7" 2>&1 | tail -4 | cut -c1-250
timeout 400 python bench.py --config w7 --steps 5 --warmup 3 --graph 0 --no-cpu-baseline --no-secondary --no-attn-isolation --no-kernel-pass > gpurun_out/r2ba_w7.json 2> gpurun_out/r2ba_w7.err; echo "w7 rc=$?"; cut -c1-300 gpurun_out/r2ba_w7.json
