#!/bin/bash
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_tc.py -x -q -k "window_attention" ) > gpurun_out/pytest_attn.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_attn.log
timeout 200 python scripts/check_tc_attn.py > gpurun_out/attn_v2.log 2>&1
MICFORMER_ATTN_V1=1 timeout 200 python scripts/check_tc_attn.py > gpurun_out/attn_v1.log 2>&1
tail -15 gpurun_out/pytest_attn.log; cat gpurun_out/attn_v2.log; tail -2 gpurun_out/attn_v1.log
