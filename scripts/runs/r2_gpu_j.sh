#!/bin/bash
mkdir -p gpurun_out
timeout 500 python scripts/loss_curve_parity.py --steps 100 --size 64 --batch 2 --gemm-mode 1 --out gpurun_out/r2j_loss_curve_mode1.json > gpurun_out/r2j_lc1.log 2>&1; tail -1 gpurun_out/r2j_lc1.log | cut -c1-500
timeout 500 python scripts/loss_curve_parity.py --steps 100 --size 64 --batch 2 --gemm-mode 0 --out gpurun_out/r2j_loss_curve_mode0.json > gpurun_out/r2j_lc0.log 2>&1; tail -1 gpurun_out/r2j_lc0.log | cut -c1-500
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2j_launches.csv python bench.py --profile-step --no-cpu-baseline > gpurun_out/r2j_launches.out 2>&1
python scripts/traffic_pass.py gpurun_out/r2j_launches.csv gpurun_out/r2j_traffic.json | head -50
