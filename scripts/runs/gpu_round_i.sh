#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-kernels gpurun_out/kernels_i.json > gpurun_out/bench_i.json 2> gpurun_out/bench_i.err
tail -6 gpurun_out/pytest_gpu.log; cut -c1-300 gpurun_out/bench_i.json
