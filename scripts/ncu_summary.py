"""Extract the judged metrics from an .ncu-rep (read on the CPU box) into a markdown table.
usage: python scripts/ncu_summary.py gpurun_out/ncu_gemm.ncu-rep > profiles/ncu_gemm.md"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader([l for l in raw.splitlines() if l.startswith('"')]))
hdr, units = rows[0], rows[1]
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"),
        ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps act %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("smsp__inst_executed.sum", "warp inst"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("launch__shared_mem_per_block_dynamic", "dyn smem"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_sb"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_inst"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch")]
idx = [(hdr.index(k), n) for k, n in want if k in hdr]
print("| " + " | ".join(f"{n} [{units[i]}]" if units[i] else n for i, n in idx) + " |")
print("|" + "---|" * len(idx))
for r in rows[2:]:
    cells = []
    for i, n in idx:
        v = r[i]
        if n == "kernel":
            v = v.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
            v = v.split("(")[0][:60]
        else:
            try:
                v = f"{float(v.replace(',', '')):.2f}"
            except ValueError:
                pass
        cells.append(v)
    print("| " + " | ".join(cells) + " |")
