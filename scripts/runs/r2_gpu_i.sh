#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity2.py tests/test_gpu_fused.py -q ) > gpurun_out/r2i_pytest_new.log 2>&1
tail -12 gpurun_out/r2i_pytest_new.log | cut -c1-300
timeout 400 python bench.py --steps 10 --warmup 3 --dump-kernels gpurun_out/r2i_kernels.json > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2i_bench.json'))
for k in ('value','ms_per_step','e2e','gpu_launches','roofline','step_roofline','cpu_baseline','attention_kernel','secondary','eager_cuda'):
    print(k, json.dumps(d.get(k))[:600])
PY
tail -3 gpurun_out/r2i_bench.err
