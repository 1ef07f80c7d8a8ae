"""Summarise an .ncu-rep (ncu --set full) into a markdown table of the metrics that matter for an
HBM-bound kernel. usage: python tools/ncu_summary.py <file.ncu-rep> "<title>" > profiles/<name>.md"""
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__block_size',
        'launch__grid_size', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__inst_executed.sum']

rep, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
print(f"# {title}\n")
print("Per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes.\n")
for r in rows[2:]:
    print("## " + r[idx['Kernel Name']].split('(')[0] + f"  (launch id {r[idx['ID']]})\n")
    print("| metric | value | unit |\n|---|---|---|")
    for w in WANT:
        if w in idx:
            print(f"| {w} | {r[idx[w]]} | {units[idx[w]]} |")
    print()
