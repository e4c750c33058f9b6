#!/bin/bash
# one gpurun call: GPU parity tests, bench, attention isolation, ncu launch list + full captures
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --dump-kernels gpurun_out/kernels.json > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python scripts/check_tc_attn.py > gpurun_out/attn.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches.csv python bench.py --profile-step --no-cpu-baseline > gpurun_out/launches.out 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:window_attn_tc_fwd -s 4 -c 1 \
    -o gpurun_out/attn_tc_fwd python scripts/check_tc_attn.py > gpurun_out/ncu_attn.out 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json | cut -c1-1500; cat gpurun_out/attn.log
