#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/grad_noise.py > gpurun_out/r2d_grad_noise.log 2>&1; tail -40 gpurun_out/r2d_grad_noise.log | cut -c1-150
( time timeout 400 python -m pytest tests -m gpu -q ) > gpurun_out/r2d_pytest_gpu.log 2>&1
tail -8 gpurun_out/r2d_pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-attn-isolation --dump-kernels gpurun_out/r2d_kernels.json > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
cut -c1-400 gpurun_out/r2d_bench.json; tail -3 gpurun_out/r2d_bench.err
