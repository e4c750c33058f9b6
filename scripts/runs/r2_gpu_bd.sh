#!/bin/bash
# round-2 session-2 evidence: GPU tests with the tcgen05 attention backward, window-7 step + kernel table, sanitizer on the
# attention kernels, ncu --set full of the backward kernel
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r2bd_pytest.log 2>&1; tail -3 gpurun_out/r2bd_pytest.log | cut -c1-200
timeout 600 python bench.py --config w7 --steps 10 --warmup 3 --graph 0 --no-cpu-baseline --no-secondary --no-attn-isolation --dump-kernels gpurun_out/r2bd_w7_kernels.json > gpurun_out/r2bd_w7.json 2> gpurun_out/r2bd_w7.err; echo "w7 rc=$?"; cut -c1-260 gpurun_out/r2bd_w7.json
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r2bd_w7_kernels.json'))
    rows = sorted(d['kernels'], key=lambda r: -r['ms_per_step'])
    print('w7 serialised kernel ms/step', round(d['total_ms_per_step'], 2))
    for r in rows[:14]: print(f"  {r['ms_per_step']:8.3f} ms {int(r['calls_per_step']):4d} x  {r['kernel']}")
except Exception as e: print('no kernel table', e)
PY
for TOOL in memcheck racecheck; do
  ( timeout 900 compute-sanitizer --tool $TOOL --print-limit 3 python -m pytest tests/test_gpu_tc.py -q -x -k "large_windows and (B0 or B4 or B5)" ) > gpurun_out/r2bd_san_${TOOL}_attn.log 2>&1
  echo "$TOOL attn: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2bd_san_${TOOL}_attn.log | tail -1)  $(grep -E ' passed| failed' gpurun_out/r2bd_san_${TOOL}_attn.log | tail -1)"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"window_attn_tc_bwd" -c 1 -o gpurun_out/r2bd_attn_bwd -f python scripts/attn_cfg4.py 1024 bwd > gpurun_out/r2bd_ncu.log 2>&1; tail -2 gpurun_out/r2bd_ncu.log
python scripts/summarize_full.py gpurun_out/r2bd_attn_bwd.ncu-rep gpurun_out/r2bd_attn_bwd.ncu.txt "round 2: tcgen05 window-attention backward, 1024 windows x 343 tokens x 96 ch x 3 heads (scripts/attn_cfg4.py 1024 bwd)" > /dev/null 2>&1; head -30 gpurun_out/r2bd_attn_bwd.ncu.txt | cut -c1-200
timeout 120 python scripts/attn_cfg4.py 4096 bwd 2>&1 | tail -2
