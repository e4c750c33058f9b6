#!/bin/bash
mkdir -p gpurun_out
timeout 240 python scripts/check_tc_attn_bwd.py > gpurun_out/r2bc_check.log 2>&1; echo "check rc=$?"; tail -9 gpurun_out/r2bc_check.log | cut -c1-220
timeout 200 python scripts/trace_attn_bwd.py > gpurun_out/r2bc_trace.txt 2>&1; echo "trace rc=$?"; cat gpurun_out/r2bc_trace.txt | cut -c1-200
