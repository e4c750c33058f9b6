#!/bin/bash
mkdir -p gpurun_out
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "rc=$?" >> gpurun_out/bench_2gpu.err
cut -c1-700 gpurun_out/bench_2gpu.json; tail -4 gpurun_out/bench_2gpu.err
