#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/trace_conv.py > gpurun_out/trace_conv.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --dump-kernels gpurun_out/kernels_q.json > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
cat gpurun_out/trace_conv.log; cut -c1-1200 gpurun_out/bench_q.json; tail -3 gpurun_out/bench_q.err
