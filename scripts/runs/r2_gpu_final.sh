#!/bin/bash
# round-2 final evidence pass (1 GPU): tests, smoke, bench (both arms), ncu launch list + DRAM bytes, ncu --set full of the
# kernels DESIGN.md discusses (summarised on the box), graph timeline
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r2f_pytest.log 2>&1; tail -3 gpurun_out/r2f_pytest.log | cut -c1-200
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2f_smoke.log 2>&1; tail -2 gpurun_out/r2f_smoke.log | cut -c1-300
timeout 900 python bench.py --dump-kernels gpurun_out/r2f_kernels.json > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; cut -c1-600 gpurun_out/r2f_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err; cut -c1-400 gpurun_out/r2f_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2f_launches.csv python bench.py --profile-step --no-cpu-baseline --no-secondary --no-attn-isolation > gpurun_out/r2f_launches.log 2>&1
python scripts/traffic_pass.py gpurun_out/r2f_launches.csv gpurun_out/r2f_traffic.json 2>&1 | tail -3
python scripts/summarize_ncu.py gpurun_out/r2f_launches.csv > gpurun_out/r2f_launches.summary.txt 2>&1; head -12 gpurun_out/r2f_launches.summary.txt | cut -c1-160
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"block_fwd_kernel|block_bwd_kernel|gemm_tc_kernel|conv3_tc_kernel|conv3_mma_bwd|ln_bwd_kernel|window_attn_tc2" -o gpurun_out/r2f_targets -f python scripts/ncu_targets.py > gpurun_out/r2f_ncu.log 2>&1; tail -2 gpurun_out/r2f_ncu.log
python scripts/summarize_full.py gpurun_out/r2f_targets.ncu-rep gpurun_out/r2f_targets.ncu.txt "round 2: one launch of each discussed kernel (scripts/ncu_targets.py)" > /dev/null 2>&1; wc -l gpurun_out/r2f_targets.ncu.txt
ls -la gpurun_out/r2f_targets.ncu-rep; du -sm gpurun_out | tail -1
python - <<'PY'
import os
p='gpurun_out/r2f_targets.ncu-rep'
if os.path.exists(p) and os.path.getsize(p) > 40e6: os.remove(p); print('removed large ncu-rep (summary kept)')
PY
timeout 300 python scripts/graph_timeline.py --size 128 --out gpurun_out/r2f_timeline_128.csv > gpurun_out/r2f_timeline_128.txt 2>&1; grep -v "  stream " gpurun_out/r2f_timeline_128.txt | head -12
