#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -k "conv3" ) > gpurun_out/pytest_conv.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_conv.log
timeout 300 python scripts/prof_conv.py 5 > gpurun_out/prof_conv.log 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-kernels gpurun_out/kernels_d.json > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err
tail -12 gpurun_out/pytest_conv.log; cat gpurun_out/prof_conv.log; tail -4 gpurun_out/pytest_gpu.log; cut -c1-300 gpurun_out/bench_d.json
