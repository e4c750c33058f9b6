#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2l_pytest.log 2>&1
tail -6 gpurun_out/r2l_pytest.log | cut -c1-300
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --no-attn-isolation --dump-kernels gpurun_out/r2l_kernels.json > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2l_bench.json'))
print('ms_per_step', d['ms_per_step'], 'launches/step', d['gpu_launches']/d['steps'])
print(json.dumps(d['roofline'])[:400])
k=json.load(open('gpurun_out/r2l_kernels.json'))
print(k['total_ms_per_step'])
for r in k['kernels'][:16]: print(f"{r['kernel']:52s} {r['ms_per_step']:.3f} {r['calls_per_step']:5.0f} {r['us_per_call']:7.1f}")
PY
tail -3 gpurun_out/r2l_bench.err
