#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/trace_tcp.py > gpurun_out/trace_tcp.log 2>&1
timeout 200 python scripts/prof_conv.py 5 > gpurun_out/prof_conv2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3_tc -c 2 -o gpurun_out/conv3_v2 python scripts/prof_conv.py 0 > gpurun_out/ncu_conv2.out 2>&1
cat gpurun_out/trace_tcp.log gpurun_out/prof_conv2.log; tail -2 gpurun_out/ncu_conv2.out
