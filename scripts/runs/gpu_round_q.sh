#!/bin/bash
mkdir -p gpurun_out
( timeout 120 python -m pytest tests/test_gpu_tc.py -x -q -k "window_attention or w7_model" ) > gpurun_out/pytest_attn.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_attn.log
timeout 60 python scripts/attn_cfg4.py 4096 > gpurun_out/attn_v2c.log 2>&1
timeout 60 python scripts/trace_conv.py > gpurun_out/trace_conv.log 2>&1
timeout 200 python bench.py --steps 10 --warmup 3 --dump-kernels gpurun_out/kernels_q.json > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
tail -5 gpurun_out/pytest_attn.log; cat gpurun_out/attn_v2c.log; cat gpurun_out/trace_conv.log; cut -c1-1500 gpurun_out/bench_q.json; tail -3 gpurun_out/bench_q.err
