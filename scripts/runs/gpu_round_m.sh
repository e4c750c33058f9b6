#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcp -s 2 -c 2 -o gpurun_out/gemm_tcp python scripts/gemm_big.py > gpurun_out/ncu_gemm.out 2>&1
tail -2 gpurun_out/ncu_gemm.out
