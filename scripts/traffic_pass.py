"""ncu launch list with DRAM bytes (one eager step of bench.py) -> per kernel family: launches, device time, measured DRAM
traffic per launch.  bench.py reads the JSON this writes (profiles/r02_traffic.json) for `roofline.traffic`.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \\
        --profile-from-start off --csv --log-file gpurun_out/x.csv python bench.py --profile-step --no-cpu-baseline
    python scripts/traffic_pass.py gpurun_out/x.csv profiles/r02_traffic.json
"""
import csv, json, sys

FAM = (("gemm_tc", "gemm"), ("gemm_kernel", "gemm"), ("colsum", "gemm"), ("conv3", "conv3"), ("_block_", "fused_block"),
       ("weight_image", "fused_block"), ("ln_fwd", "layernorm"), ("ln_bwd", "layernorm"), ("window_attn", "window_attn"),
       ("deform", "deform"), ("offset_head", "offset_head"), ("dice", "dice"), ("adam", "adam"), ("block_permute", "block_permute"))


def fam_of(name):
    for key, f in FAM:
        if key in name:
            return f
    return "torch/other"


rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
i_name, i_metric, i_unit, i_val, i_id = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value", "ID"))
per = {}
for r in rows[1:]:
    d = per.setdefault(r[i_id], {"name": r[i_name]})
    v = float(r[i_val].replace(",", ""))
    u = r[i_unit]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0,
             "msecond": 1e3, "second": 1e6}.get(u, 1)
    d[r[i_metric]] = v * scale
fams = {}
for d in per.values():
    f = fams.setdefault(fam_of(d["name"]), {"launches": 0, "us": 0.0, "dram_bytes": 0.0})
    f["launches"] += 1
    f["us"] += d.get("gpu__time_duration.sum", 0.0)
    f["dram_bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
tot = sum(f["us"] for f in fams.values())
out = {"source": sys.argv[1], "launches": len(per), "total_us": round(tot, 1), "families": {}}
for k, f in sorted(fams.items(), key=lambda kv: -kv[1]["us"]):
    out["families"][k] = {"launches": f["launches"], "us": round(f["us"], 1), "share": round(f["us"] / tot, 4),
                          "dram_bytes_per_launch": int(f["dram_bytes"] / f["launches"]), "dram_gb_per_step": round(f["dram_bytes"] / 1e9, 3)}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1)[:1500])
