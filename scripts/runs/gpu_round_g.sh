#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_tc.py -x -q -k "window_attention" ) > gpurun_out/pytest_attn.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_attn.log
timeout 200 python scripts/attn_cfg4.py 4096 > gpurun_out/attn_v2b.log 2>&1
for thr in 0 1024 8192; do
  echo "== small_m=$thr" >> gpurun_out/tf32_small.log
  MICFORMER_TF32_FWD_SMALL_M=$thr timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -s -k "whole_model or w7_model" 2>&1 | grep -E "rel err|passed|failed" >> gpurun_out/tf32_small.log
  MICFORMER_TF32_FWD_SMALL_M=$thr timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-pass 2>/dev/null | cut -c1-200 >> gpurun_out/tf32_small.log
done
tail -5 gpurun_out/pytest_attn.log; cat gpurun_out/attn_v2b.log gpurun_out/tf32_small.log
