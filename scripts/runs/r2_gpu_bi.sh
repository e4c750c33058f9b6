#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity2.py -q -x -k "window_attention or w7 or 4096 or config4 or attn" 2>&1 | tail -3 | cut -c1-200
timeout 200 python scripts/attn_cfg4.py 4096 2>&1 | tail -1
MICFORMER_ATTN_ISSUERS=1 timeout 200 python scripts/attn_cfg4.py 4096 2>&1 | tail -1
timeout 200 python scripts/check_tc_attn.py 2>&1 | tail -7 | cut -c1-200
