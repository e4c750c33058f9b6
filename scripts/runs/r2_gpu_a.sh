#!/bin/bash
# round 2, call A: sanity of the round-1 state + the experiments round 1 left unrun
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
( timeout 120 scripts/ubench/ubench ) > gpurun_out/r2a_ubench.txt 2>&1
echo "ubench rc=$?" >> gpurun_out/r2a_ubench.txt
( timeout 100 python scripts/attn_cfg4.py 4096 ) > gpurun_out/r2a_attn_issuers1.log 2>&1
( MICFORMER_ATTN_ISSUERS=2 timeout 100 python scripts/attn_cfg4.py 4096 ) > gpurun_out/r2a_attn_issuers2.log 2>&1
( MICFORMER_ATTN_ISSUERS=2 timeout 200 python -m pytest tests -m gpu -x -q -k "window_attention or tc_window" ) > gpurun_out/r2a_pytest_attn2.log 2>&1
( time timeout 300 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2a_pytest_gpu.log 2>&1
tail -3 gpurun_out/r2a_pytest_gpu.log; tail -3 gpurun_out/r2a_pytest_attn2.log; cat gpurun_out/r2a_attn_issuers1.log gpurun_out/r2a_attn_issuers2.log; tail -40 gpurun_out/r2a_ubench.txt
