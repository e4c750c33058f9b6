#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_parity2.py -q -x -k "conv3 or whole_model or full_size or known_answers or tiny64 or block_vs or odd or w7" 2>&1 | tail -4 | cut -c1-250
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --no-attn-isolation --no-kernel-pass > gpurun_out/r2bk_bench_$name.json 2> gpurun_out/r2bk_bench_$name.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2bk_bench_$name.json')); print('$name', round(d['ms_per_step'],3), d['gpu_launches'])
except Exception as e: print('$name failed', e)
PY
}
run new MICFORMER_CONV_BW16_OLD=0
run old MICFORMER_CONV_BW16_OLD=1
run new2 MICFORMER_CONV_BW16_OLD=0
python - <<'PY'
import sys, torch
sys.path.insert(0, '.')
from micformer_b200 import ops, _native as N
N.set_gemm_mode(1)
def t(B, D, C0, C1, Co, ncdhw, tag):
    dy = torch.randn((B, Co, D, D, D) if ncdhw else (B, D, D, D, Co), device='cuda'); x0 = torch.randn(B, D, D, D, C0, device='cuda')
    x1 = torch.randn(B, D, D, D, C1, device='cuda') if C1 else None
    dw = torch.zeros(27, C0 + C1, Co, device='cuda'); db = torch.zeros(Co, device='cuda')
    for _ in range(2): ops.conv3_bwd_weight(dy, x0, x1, dw, db, B, (D, D, D), Co, ncdhw)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ops.conv3_bwd_weight(dy, x0, x1, dw, db, B, (D, D, D), Co, ncdhw)
    e1.record(); torch.cuda.synchronize()
    print('%s: %.1f us' % (tag, e0.elapsed_time(e1) / 5 * 1e3))
t(2, 128, 24, 0, 8, True, 'out_conv 2x128^3 24->8')
t(2, 32, 48, 48, 16, False, 'conv_offset stage 0 (2x32^3, 96->16)')
t(2, 16, 96, 96, 16, False, 'conv_offset stage 1 (2x16^3, 192->16)')
t(2, 8, 192, 192, 16, False, 'conv_offset stage 2 (2x8^3, 384->16)')
t(2, 4, 384, 384, 16, False, 'conv_offset stage 3 (2x4^3, 768->16)')
PY
