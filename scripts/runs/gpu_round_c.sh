#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/prof_conv.py 5 > gpurun_out/prof_conv.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3_ -c 9 -o gpurun_out/conv3_full python scripts/prof_conv.py 0 > gpurun_out/ncu_conv.out 2>&1
cat gpurun_out/prof_conv.log; tail -3 gpurun_out/ncu_conv.out
