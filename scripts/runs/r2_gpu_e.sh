#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/determinism.py > gpurun_out/r2e_determinism.log 2>&1; tail -40 gpurun_out/r2e_determinism.log | cut -c1-150
( time timeout 400 python -m pytest tests -m gpu -q ) > gpurun_out/r2e_pytest_gpu.log 2>&1
tail -8 gpurun_out/r2e_pytest_gpu.log
