#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-pass 2> gpurun_out/bench_n.err | cut -c1-260 > gpurun_out/bench_n_pdl.json
MICFORMER_PDL=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-pass 2>/dev/null | cut -c1-260 > gpurun_out/bench_n_nopdl.json
cat gpurun_out/bench_n_pdl.json gpurun_out/bench_n_nopdl.json; tail -5 gpurun_out/bench_n.err
