#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_optim.py -x -q ) > gpurun_out/pytest_arena.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_arena.log
tail -15 gpurun_out/pytest_arena.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-pass 2> gpurun_out/bench_o.err | cut -c1-260 > gpurun_out/bench_o_arena.json
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-pass --arena 0 2>/dev/null | cut -c1-260 > gpurun_out/bench_o_noarena.json
cat gpurun_out/bench_o_arena.json gpurun_out/bench_o_noarena.json; tail -5 gpurun_out/bench_o.err
