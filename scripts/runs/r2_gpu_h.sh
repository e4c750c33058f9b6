#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_fused.py -x -q ) > gpurun_out/r2h_fused.log 2>&1; tail -5 gpurun_out/r2h_fused.log | cut -c1-250
timeout 100 python scripts/time_fused.py 2>&1 | tail -5
