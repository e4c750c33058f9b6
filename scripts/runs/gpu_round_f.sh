#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:window_attn_tc2 -s 2 -c 1 -o gpurun_out/attn_tc2_fwd python scripts/attn_cfg4.py 1024 > gpurun_out/ncu_attn2.out 2>&1
timeout 200 python scripts/trace_tc.py > gpurun_out/trace_tc.log 2>&1
tail -3 gpurun_out/ncu_attn2.out; cat gpurun_out/trace_tc.log
