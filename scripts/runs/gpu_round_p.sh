#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu_arena.json 2> gpurun_out/bench_2gpu_arena.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --arena 0 > gpurun_out/bench_2gpu_noarena.json 2> gpurun_out/bench_2gpu_noarena.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-pass 2>/dev/null | cut -c1-260 > gpurun_out/bench_p_1gpu.json
cut -c1-330 gpurun_out/bench_2gpu_arena.json; cut -c1-330 gpurun_out/bench_2gpu_noarena.json; cat gpurun_out/bench_p_1gpu.json; tail -3 gpurun_out/bench_2gpu_arena.err
