#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_fused.py -x -q ) > gpurun_out/r2f_fused.log 2>&1; tail -25 gpurun_out/r2f_fused.log | cut -c1-250
( time timeout 400 python -m pytest tests -m gpu -q ) > gpurun_out/r2f_pytest_gpu.log 2>&1
tail -8 gpurun_out/r2f_pytest_gpu.log | cut -c1-250
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-attn-isolation --dump-kernels gpurun_out/r2f_kernels.json > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
cut -c1-330 gpurun_out/r2f_bench.json; tail -3 gpurun_out/r2f_bench.err
