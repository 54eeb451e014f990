"""Summarise one kernel of an ncu report: headline metrics, stall ratios, per-source-line instruction / sample shares.
usage: ncu_kernel.py report.ncu-rep kernel-regex [top]"""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--kernel-name", "regex:" + pat, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "launch__shared_mem_per_block_dynamic"]
want += [h for h in hdr if "issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h]
d = dict(zip(hdr, rows[2]))
for w in want:
    if w in d and d[w] not in ("0", "0.000000"):
        print(w.replace("smsp__average_warps_issue_stalled_", "stall_").replace("_per_issue_active.ratio", ""), "=", d[w])
src = subprocess.run(["ncu", "-i", rep, "--kernel-name", "regex:" + pat, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]; ii = hdr.index("Instructions Executed"); sm = hdr.index("# Samples")
data = []
for r in rows[hi + 1:]:
    if len(r) <= ii or r[2] != "-":
        continue
    try:
        data.append((int(float(r[ii] or 0)), int(float(r[sm] or 0)), r[0], r[1].strip()[:105]))
    except ValueError:
        pass
    if data and data[-1][2] == "1" and len(data) > 50 and False:
        break
tot = sum(x[0] for x in data); ts = sum(x[1] for x in data)
print("total warp-instructions", tot, "samples", ts)
for x in sorted(data, key=lambda x: -x[1])[:top]:
    print(f"{100*x[0]/tot:5.1f}% inst {100*x[1]/ts:5.1f}% smp  L{x[2]:>4} {x[3]}")
