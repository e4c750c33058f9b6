#!/bin/bash
mkdir -p gpurun_out
( timeout 180 python -m pytest tests/test_gpu_fused.py -x -q ) > gpurun_out/r2b_fused.log 2>&1
echo "rc=$?" >> gpurun_out/r2b_fused.log
tail -30 gpurun_out/r2b_fused.log
( timeout 120 python scripts/time_fused.py ) 2>&1 | tail -4
