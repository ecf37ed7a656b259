"""Pick the judged metrics out of `ncu --page raw --csv` (stdin) -> one block per launch."""
import csv, sys
KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
rows = list(csv.reader(sys.stdin))
hdr = None
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr = i
        break
if hdr is None:
    sys.exit("no header")
names, units = rows[hdr], rows[hdr + 1]
for r in rows[hdr + 2:]:
    if len(r) != len(names):
        continue
    d = dict(zip(names, r))
    u = dict(zip(names, units))
    print("-" * 100)
    for k in KEYS:
        if k in d:
            print(f"{k:85s} {d[k]} {u.get(k, '')}")
    try:
        rd = float(d["dram__bytes_read.sum"].replace(",", "")); wr = float(d["dram__bytes_write.sum"].replace(",", ""))
        t = float(d["gpu__time_duration.sum"].replace(",", ""))
        print(f"{'[derived] dram traffic (read+write) in units of ' + u.get('dram__bytes_read.sum', ''):85s} {rd + wr:.3f}; duration unit {u.get('gpu__time_duration.sum', '')} = {t}")
    except Exception:
        pass
