#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -k "linear or epilogues" ) > gpurun_out/pytest_gemm.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gemm.log
tail -15 gpurun_out/pytest_gemm.log
if grep -q "rc=0" gpurun_out/pytest_gemm.log; then
  timeout 200 python scripts/trace_tcp.py 2>&1 | grep -v "tile [1-4]:" > gpurun_out/trace_tcp2.log
  ( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-kernels gpurun_out/kernels_l.json > gpurun_out/bench_l.json 2> gpurun_out/bench_l.err
  cat gpurun_out/trace_tcp2.log; tail -6 gpurun_out/pytest_gpu.log; cut -c1-300 gpurun_out/bench_l.json
fi
