#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/trace_attn_bwd.py > gpurun_out/r2bb_trace.txt 2>&1; echo "trace rc=$?"; cat gpurun_out/r2bb_trace.txt | cut -c1-200
