"""Print the key metrics of every kernel in an .ncu-rep (raw page): python scripts/ncu_keys.py file.ncu-rep [regex]"""
import csv, re, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else
                 r"^(Kernel Name|gpu__time_duration.sum|dram__bytes_read.sum|dram__bytes_write.sum|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|"
                 r"sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active|sm__inst_executed_pipe_tensor.*sum$|sm__warps_active.avg.pct_of_peak_sustained_active|launch__registers_per_thread|"
                 r"launch__grid_size|launch__block_size|launch__occupancy_limit_.*|sm__throughput.avg.pct_of_peak_sustained_elapsed|lts__t_bytes.sum|"
                 r"lts__t_sector_hit_rate.pct|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum|smsp__average_warp.*_per_issue_active.*|smsp__issue_active.avg.pct.*|"
                 r"sm__pipe_tensor.*realtime.*pct.*|launch__waves_per_multiprocessor|sm__cycles_active.avg)$")
for r in rows[2:]:
    print("-" * 100)
    for h, u, v in zip(hdr, units, r):
        if pat.search(h):
            print(f"{h:90s} {v} {u}")
