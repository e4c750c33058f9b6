#!/bin/bash
# round-end evidence: GPU parity tests, both bench arms, ncu launch list of one eager step, full captures of the
# attention kernel and of the step's LayerNorm / small-window attention kernels
mkdir -p gpurun_out
( time timeout 200 python -m pytest tests -m gpu -x -q ) > gpurun_out/final_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/final_pytest_gpu.log
timeout 200 python bench.py --steps 10 --warmup 3 --dump-kernels gpurun_out/final_kernels.json > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
timeout 100 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2>/dev/null
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/final_launches.csv python bench.py --profile-step --no-cpu-baseline > gpurun_out/final_launches.out 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:window_attn_tc2 -s 2 -c 1 \
    -o gpurun_out/final_attn_tc2 python scripts/attn_cfg4.py 1024 > gpurun_out/final_ncu_attn.out 2>&1
timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k "regex:ln_bwd_kernel|ln_fwd_kernel|window_attn_bwd_kernel|window_attn_fwd_kernel" -c 4 \
    -o gpurun_out/final_ln_attn python bench.py --profile-step --no-cpu-baseline > gpurun_out/final_ncu_ln.out 2>&1
tail -4 gpurun_out/final_pytest_gpu.log; cut -c1-400 gpurun_out/final_bench.json; cut -c1-300 gpurun_out/final_bench_reference.json; tail -2 gpurun_out/final_ncu_attn.out; tail -2 gpurun_out/final_ncu_ln.out
