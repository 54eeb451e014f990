"""Key metrics per kernel from `ncu -i X.ncu-rep --page raw --csv`.  usage: ncu_summary.py raw.csv"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "launch__shared_mem_per_block_dynamic", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    m = re.search(r"cmpc_\w+", d["Kernel Name"])
    print("=== %s" % (m.group(0) if m else d["Kernel Name"]))
    for k in want:
        if k in d:
            print("%s = %s" % (k, d[k]))
    st = sorted(((float(d[h]), h) for h in stall if d[h] not in ("", "n/a")), reverse=True)[:6]
    for v, h in st:
        print("stall_%s (warps per issue) = %.3f" % (h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v))
    print()
