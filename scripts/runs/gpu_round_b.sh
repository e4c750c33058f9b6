#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -k "conv3" ) > gpurun_out/pytest_conv.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_conv.log
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --dump-kernels gpurun_out/kernels_b.json > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
tail -4 gpurun_out/pytest_conv.log; tail -4 gpurun_out/pytest_gpu.log; cut -c1-400 gpurun_out/bench_b.json
