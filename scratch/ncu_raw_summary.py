"""Per-kernel means of selected metrics from `ncu -i X.ncu-rep --page raw --csv` output.
python scratch/ncu_raw_summary.py raw.csv [raw2.csv ...] > profiles/....txt"""
import csv, sys, collections
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum.per_cycle_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    col = {n: i for i, n in enumerate(names)}
    agg = collections.OrderedDict()
    for r in rows[hdr + 2:]:
        if len(r) < len(names):
            continue
        k = r[col["Kernel Name"]]
        agg.setdefault(k, []).append(r)
    print(f"## {path}")
    for k, rs in agg.items():
        short = k.split("(")[0]
        print(f"### {short}   ({len(rs)} launches captured; grid {rs[0][col['Grid Size']]} x block {rs[0][col['Block Size']]})")
        for m in KEYS:
            if m not in col:
                continue
            vals = []
            for r in rs:
                try:
                    vals.append(float(r[col[m]].replace(",", "")))
                except ValueError:
                    pass
            if vals:
                print(f"    {m:90s} {sum(vals) / len(vals):14.3f} {units[col[m]]}")
        try:
            t = sum(float(r[col["gpu__time_duration.sum"]].replace(",", "")) for r in rs) / len(rs)
            b = sum(float(r[col["dram__bytes_read.sum"]].replace(",", "")) + float(r[col["dram__bytes_write.sum"]].replace(",", "")) for r in rs) / len(rs)
            print(f"    => DRAM traffic per launch {b:.4g} {units[col['dram__bytes_read.sum']]} in {t:.4g} {units[col['gpu__time_duration.sum']]}")
        except Exception:
            pass
