"""Per-kernel launch counts, mean device time and shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: launch_shares.py launches.csv"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
t = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0].replace("<unnamed>::", "")
    t.setdefault(name, []).append(float(r[14]) / 1e3)
solve = [k for k in t if any(s in k for s in ("assemble", "invert", "lpt_order", "dual", "condense", "solve_kernel"))]
tot = sum(sum(t[k]) for k in solve)
print("# per-launch times are cold-cache and serialised (ncu): compare shares, not absolutes")
for k, v in t.items():
    share = "  share of the solve kernels %5.1f%%" % (100 * sum(v) / tot) if k in solve else ""
    print("%-48s launches %4d  mean %8.1f us%s" % (k[:48], len(v), sum(v) / len(v), share))
