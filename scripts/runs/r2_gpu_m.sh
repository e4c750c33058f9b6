#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -4 | tee gpurun_out/r2m_multi.log
for ov in 1 0; do
MICFORMER_OVERLAP=$ov timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2m_bench2_ov$ov.json 2> gpurun_out/r2m_bench2_ov$ov.err
python -c "
import json; d=json.load(open('gpurun_out/r2m_bench2_ov$ov.json')); print('overlap=$ov', d['n_gpus'], d['ms_per_step'], d['value'])"
done
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary --no-attn-isolation --no-kernel-pass > gpurun_out/r2m_bench1.json 2> gpurun_out/r2m_bench1.err
python -c "
import json; d=json.load(open('gpurun_out/r2m_bench1.json')); print('1 gpu', d['ms_per_step'], d['value'])"
