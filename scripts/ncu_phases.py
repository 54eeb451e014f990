"""Share of executed instructions and stall samples per kernel phase (from an ncu --set full report)."""
import csv, re, subprocess, sys
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
cur = None; data = []
hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if "Instructions Executed" in r:
        hdr = r; ii = hdr.index("Instructions Executed"); sm = hdr.index("# Samples"); continue
    if hdr is None or len(r) <= ii or r[2] != "-":
        continue
    try:
        data.append((cur, int(r[0]), int(float(r[ii] or 0)), int(float(r[sm] or 0)), r[1]))
    except ValueError:
        pass
def marks(path, pats):
    out = {}
    for n, line in enumerate(open(path), 1):
        for name, pat in pats.items():
            if name not in out and pat in line:
                out[name] = n
    return out
root = __file__.rsplit("/scripts", 1)[0] + "/quad-periodic-mpc_b200/csrc/"
mk = marks(root + "cmpc_kernels.cu", {"kernel": "cmpc_solve_kernel(const __grid_constant__", "est": "---- 0. periodic", "A1": "---- A1.",
                                      "C": "---- C. gradient", "D": "---- D. K <- H^-1", "x0": "// x = -H^-1 g", "E": "---- E. Goldfarb",
                                      "obj": "---- objective", "F": "---- F. outputs"})
ms = marks(root + "cmpc_sweep.cuh", {"hbuild": "H into shared memory", "load": "double A[TM][TN];", "loop": "for (int k = 0; k < n; k++)", "store": "-swept = (scaled H)^-1"})
def phase(f, l):
    if f == "cmpc_sweep.cuh":
        if l < ms["hbuild"]: return "sweep:rcp-helper"
        if l < ms["load"]: return "D1 Hessian assembly"
        if l < ms["loop"]: return "D2 tile load/diag init"
        if l < ms["store"]: return "D3 pivot loop"
        return "D4 K store"
    if f == "cmpc_adapt.cuh": return "0 estimator"
    if f == "cmpc_kernels.cu":
        if l < mk["kernel"]: return "helpers (reductions, cons_of, psym)"
        if l < mk["A1"]: return "prologue / record wait"
        if l < mk["C"]: return "A-B state, W, e, aggregates"
        if l < mk["x0"]: return "C gradient"
        if l < mk["E"]: return "x0 = -Kg, slacks"
        if l < mk["obj"]: return "E active-set iterations"
        if l < mk["F"]: return "objective"
        return "F outputs"
    return "other:" + str(f)
agg = {}
for f, l, ins, smp, _ in data:
    p = phase(f, l); a = agg.setdefault(p, [0, 0]); a[0] += ins; a[1] += smp
ti = sum(a[0] for a in agg.values()); tsm = sum(a[1] for a in agg.values())
print(f"{'phase':40s} {'inst%':>7s} {'samples%':>9s}")
for p, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{p:40s} {100*a[0]/ti:7.1f} {100*a[1]/tsm:9.1f}")
print("total warp-instructions", ti, "samples", tsm)
