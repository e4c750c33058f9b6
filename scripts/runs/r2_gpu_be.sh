#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,launch__grid_size,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"offset_head|deform_sample" --csv --log-file gpurun_out/r2be_w7_offs.csv python bench.py --config w7 --steps 1 --warmup 1 --graph 0 --no-cpu-baseline --no-secondary --no-attn-isolation --no-kernel-pass > gpurun_out/r2be.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/r2be_w7_offs.csv')) if len(r) > 10]
hdr = rows[0]; ik = hdr.index('Kernel Name'); im = hdr.index('Metric Name'); iv = hdr.index('Metric Value'); iid = hdr.index('ID')
d = {}
for r in rows[1:]:
    d.setdefault((r[iid], r[ik][:40]), {})[r[im]] = r[iv]
n = 0
for (i, k), m in d.items():
    if 'offset_head_bwd' in k and n < 60:
        n += 1
        print(i, k, m.get('gpu__time_duration.sum'), 'grid', m.get('launch__grid_size'), 'rd', m.get('dram__bytes_read.sum'), 'wr', m.get('dram__bytes_write.sum'))
PY
