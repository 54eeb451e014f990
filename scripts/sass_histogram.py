"""Per-kernel counts of the SASS mnemonics that prove what the kernels are built from (cuobjdump -sass of the shipped
library): FP64 tensor (DMMA), FP64 FMA (DFMA), bulk async copy (UBLKCP = cp.async.bulk), mbarrier (SYNCS), warp
reductions (REDUX), setmaxnreg (USETMAXREG), named barriers (BAR.ARV / BAR.SYNC).  usage: sass_histogram.py lib.so"""
import re, subprocess, sys, collections
lib = sys.argv[1]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
pats = collections.OrderedDict([("DMMA", r"\bDMMA\."), ("DFMA", r"\bDFMA\b"), ("DMUL/DADD", r"\b(DMUL|DADD)\b"), ("UBLKCP", r"\bUBLKCP"),
                                ("SYNCS (mbarrier)", r"\bSYNCS"), ("REDUX / CREDUX", r"\bC?REDUX"), ("USETMAXREG", r"USETMAXREG"),
                                ("BAR.ARV", r"\bBAR\.ARV"), ("BAR.SYNC", r"\bBAR\.SYNC"), ("SHFL", r"\bSHFL\."),
                                ("LDS/STS", r"\b(LDS|STS)\b"), ("MUFU.RCP64H", r"MUFU\.RCP64H"),
                                ("UTCxMMA / LDTM / UTMALDG (tcgen05, TMA tensor)", r"\b(UTC\w*MMA|LDTM|UTMALDG)")])
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
arch = set(re.findall(r"arch = (sm_\w+)", sass))
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(anonymous namespace\)::", "", cur).split("(")[0]
        counts[cur] = collections.Counter()
        continue
    if cur is None or "/*" not in line:
        continue
    total[cur] += 1
    for k, p in pats.items():
        if re.search(p, line):
            counts[cur][k] += 1
print("# cuobjdump -sass %s   (architectures: %s)" % (lib.split("/")[-1], ", ".join(sorted(arch))))
print("# tcgen05 has no FP64 type and the staging is 1-D bulk copies, so UTC*MMA / LDTM / UTMALDG are expected to be absent")
for fn, c in counts.items():
    if total[fn] < 50:
        continue
    print("%-72s %6d instr  %s" % (fn[:72], total[fn], "  ".join("%s %d" % (k, c[k]) for k in pats if c[k])))
