#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -q -x -k "conv3 or whole_model or full_size or known_answers or tiny64" 2>&1 | tail -4 | cut -c1-250
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --no-attn-isolation --no-kernel-pass > gpurun_out/r2bj_bench_$name.json 2> gpurun_out/r2bj_bench_$name.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2bj_bench_$name.json')); print('$name', round(d['ms_per_step'],3), d['gpu_launches'])
except Exception as e: print('$name failed', e)
PY
}
run new MICFORMER_CONV_BW8_OLD=0
run old MICFORMER_CONV_BW8_OLD=1
run new2 MICFORMER_CONV_BW8_OLD=0
python - <<'PY'
import sys, torch
sys.path.insert(0, '.')
from micformer_b200 import ops, _native as N
N.set_gemm_mode(1)
B, D = 2, 128
dy = torch.randn(B, 8, D, D, D, device='cuda'); x = torch.randn(B, D, D, D, 24, device='cuda')
dw = torch.zeros(27, 24, 8, device='cuda'); db = torch.zeros(8, device='cuda')
for _ in range(2): ops.conv3_bwd_weight(dy, x, None, dw, db, B, (D, D, D), 8, True)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): ops.conv3_bwd_weight(dy, x, None, dw, db, B, (D, D, D), 8, True)
e1.record(); torch.cuda.synchronize()
print('out_conv bwd-weight (2 x 128^3, 24 -> 8): %.3f ms' % (e0.elapsed_time(e1) / 5))
PY
