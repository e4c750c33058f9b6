#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -k "linear or epilogues" ) > gpurun_out/pytest_gemm.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gemm.log
timeout 200 python scripts/trace_tc.py > gpurun_out/trace_tc2.log 2>&1
tail -4 gpurun_out/pytest_gemm.log; grep -E "^---|^mean" gpurun_out/trace_tc2.log
