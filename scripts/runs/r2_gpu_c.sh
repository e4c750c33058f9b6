#!/bin/bash
mkdir -p gpurun_out
for T in 18944 37888 75776 151552; do timeout 100 python scripts/time_fused.py $T 2>&1 | tail -2; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp_block -s 6 -c 2 -o gpurun_out/r2c_mlp python scripts/time_fused.py 65536 > gpurun_out/r2c_ncu.log 2>&1
tail -3 gpurun_out/r2c_ncu.log
