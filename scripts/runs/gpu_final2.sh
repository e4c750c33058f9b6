#!/bin/bash
mkdir -p gpurun_out
( time timeout 200 python -m pytest tests -m gpu -x -q ) > gpurun_out/final2_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/final2_pytest_gpu.log
timeout 200 python bench.py --steps 10 --warmup 3 --dump-kernels gpurun_out/final2_kernels.json > gpurun_out/final2_bench.json 2> gpurun_out/final2_bench.err
timeout 60 python scripts/attn_cfg4.py 4096 > gpurun_out/final2_attn.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/final2_launches.csv python bench.py --profile-step --no-cpu-baseline > gpurun_out/final2_launches.out 2>&1
tail -4 gpurun_out/final2_pytest_gpu.log; cut -c1-330 gpurun_out/final2_bench.json; cat gpurun_out/final2_attn.log
