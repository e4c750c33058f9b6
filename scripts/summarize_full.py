"""Summarise an `ncu --set full --import-source on` report into a small text file for profiles/:
    python scripts/summarize_full.py gpurun_out/x.ncu-rep profiles/r01_x.ncu.txt ["title"]
Per captured launch: duration, DRAM bytes/throughput, tensor-pipe and issue utilisation, occupancy limiters, the warp
stall mix and the ten source lines with the most stall samples (needs -lineinfo, which the build passes)."""
import csv, re, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
        ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (of active cycles)"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers/thread"),
        ("launch__occupancy_limit_registers", "CTAs/SM limit: registers"), ("launch__occupancy_limit_shared_mem", "CTAs/SM limit: smem"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts")]
lines = [f"# {title}", f"# source: {rep} (ncu --set full --clock-control none --import-source on; cold-cache, serialised replays)", ""]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    lines.append(f"## {name[:110]}")
    for k, label in KEYS:
        if k in idx:
            lines.append(f"  {label:46s} {r[idx[k]]} {units[idx[k]]}")
    stalls = []
    for h, i in idx.items():
        m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active.ratio", h)
        if m:
            try: stalls.append((float(r[i]), m.group(1)))
            except ValueError: pass
    stalls.sort(reverse=True)
    lines.append("  warp stall mix (stalled warps per issued instruction): " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:6]))
    lines.append("")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
srows = list(csv.reader(src.splitlines()))
starts = [i for i, r in enumerate(srows) if r and r[0] == "Function Name"]
for kk, st in enumerate(starts):
    sub = srows[st:starts[kk + 1] if kk + 1 < len(starts) else len(srows)]
    his = [i for i, r in enumerate(sub) if any(c.startswith("Warp Stall Sampling (All") for c in r)]
    if not his:
        continue
    h = sub[his[0]]
    ia = [i for i, c in enumerate(h) if c.startswith("Warp Stall Sampling (All")][0]
    ie = h.index("Instructions Executed")
    ls = []
    for r in sub[his[0] + 1:]:
        if r[0].strip().isdigit():
            try: ls.append((int(r[ia] or 0), int(r[0]), r[1].strip(), int(r[ie] or 0)))
            except ValueError: pass
    tot = sum(l[0] for l in ls)
    if tot < 200:
        continue
    lines.append(f"## stall samples by source line: {sub[0][1][:90]}  ({tot} samples)")
    for s_, ln, code, ex in sorted(ls, key=lambda d: -d[0])[:10]:
        lines.append(f"  {100 * s_ / tot:5.1f}%  warp-instr {ex:>10d}  L{ln}: {code[:100]}")
    lines.append("")
open(out, "w").write("\n".join(lines) + "\n")
print(out, len(lines), "lines")
