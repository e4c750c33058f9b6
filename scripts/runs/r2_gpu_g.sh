#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:block_ -s 8 -c 2 -o gpurun_out/r2g_mlp python scripts/time_fused.py 65536 > gpurun_out/r2g_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_block_ -s 8 -c 1 -o gpurun_out/r2g_attn_f python scripts/time_fused.py 65536 >> gpurun_out/r2g_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_block_bwd -s 8 -c 1 -o gpurun_out/r2g_attn_b python scripts/time_fused.py 65536 >> gpurun_out/r2g_ncu.log 2>&1
tail -2 gpurun_out/r2g_ncu.log; ls -la gpurun_out/*.ncu-rep
