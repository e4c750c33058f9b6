#!/bin/bash
# colsum fold with shuffle reduction; graph timelines (CUPTI through torch.profiler) at 128^3 and 64^3
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -q -x -k "linear or epilogues" 2>&1 | tail -3 | cut -c1-250
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --no-attn-isolation --no-kernel-pass $EXTRA > gpurun_out/r2r_bench_$name.json 2> gpurun_out/r2r_bench_$name.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2r_bench_$name.json')); print('$name', round(d['ms_per_step'],3), d['gpu_launches'])
except Exception as e: print('$name failed', e)
PY
}
run nofold MICFORMER_FOLD_COLSUM=0
run fold MICFORMER_FOLD_COLSUM=1
timeout 300 python scripts/graph_timeline.py --size 128 --out gpurun_out/r2r_timeline_128.csv > gpurun_out/r2r_timeline_128.txt 2>&1; cat gpurun_out/r2r_timeline_128.txt | cut -c1-200
timeout 300 python scripts/graph_timeline.py --size 64 --out gpurun_out/r2r_timeline_64.csv > gpurun_out/r2r_timeline_64.txt 2>&1; cat gpurun_out/r2r_timeline_64.txt | cut -c1-200
