"""Summarise an .ncu-rep (read here, no GPU needed) into a small markdown table for profiles/.
Usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/NAME.md"""
import csv, subprocess, sys, io

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 read sectors (from L1)"),
    ("launch__registers_per_thread", "registers/thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"), ("launch__shared_mem_per_block_static", "static smem/block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__maximum_warps_per_active_cycle_pct", "theoretical occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp instruction"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard (warps/issue)"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch_resolving"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall membar"),
    ("sm__cycles_elapsed.max", "SM cycles"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    names = [r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "") for r in data]
    print(f"# ncu --set full summary of `{rep.split('/')[-1]}`\n")
    print("| metric | unit | " + " | ".join(f"#{i} {n}" for i, n in enumerate(names)) + " |")
    print("|---|---|" + "---|" * len(names))
    for key, label in METRICS:
        if key not in hdr:
            continue
        i = hdr.index(key)
        vals = []
        for r in data:
            try:
                v = float(r[i].replace(",", "")); vals.append(f"{v:,.2f}" if abs(v) < 1e6 else f"{v:,.0f}")
            except ValueError:
                vals.append(r[i])
        print(f"| {label} (`{key}`) | {units[i]} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
