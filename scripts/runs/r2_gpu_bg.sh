#!/bin/bash
# sanitizer on the large-window attention kernels (forward + tcgen05 backward), then stream-priority A/B on the train step
mkdir -p gpurun_out
for TOOL in memcheck racecheck; do
  ( timeout 900 compute-sanitizer --tool $TOOL --print-limit 3 python -m pytest tests/test_gpu_tc.py -q -x -k "large_windows and (1-pd0 or 1-pd4 or 2-pd5)" ) > gpurun_out/r2bg_san_${TOOL}_attn.log 2>&1
  echo "$TOOL attn: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2bg_san_${TOOL}_attn.log | tail -1)  $(grep -E ' passed| failed' gpurun_out/r2bg_san_${TOOL}_attn.log | tail -1)"
done
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --no-attn-isolation --no-kernel-pass > gpurun_out/r2bg_bench_$name.json 2> gpurun_out/r2bg_bench_$name.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2bg_bench_$name.json')); print('$name', round(d['ms_per_step'],3), d['gpu_launches'])
except Exception as e: print('$name failed', e)
PY
}
run p0 MICFORMER_STREAM_PRIO=0
run p1 MICFORMER_STREAM_PRIO=1
run p0b MICFORMER_STREAM_PRIO=0
run p1b MICFORMER_STREAM_PRIO=1
