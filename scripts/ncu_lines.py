"""Summarise an ncu report: headline metrics and the per-source-line instruction / stall-sample shares."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[-1]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "launch__shared_mem_per_block_dynamic", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
for h, u, v in zip(hdr, units, vals):
    if h in want:
        print(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]; ii = hdr.index("Instructions Executed"); sm = hdr.index("# Samples")
data = []
for r in rows[hi + 1:]:
    if len(r) <= ii or r[2] != "-":
        continue
    try:
        data.append((int(float(r[ii] or 0)), int(float(r[sm] or 0)), r[0], r[1].strip()[:100]))
    except ValueError:
        pass
tot = sum(d[0] for d in data); ts = sum(d[1] for d in data)
print("total warp-instructions", tot, "samples", ts)
for d in sorted(data, key=lambda x: -x[1])[:top]:
    print(f"{100*d[0]/tot:5.1f}% inst {100*d[1]/ts:5.1f}% smp  L{d[2]:>4} {d[3]}")
