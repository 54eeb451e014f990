"""Stall samples of one kernel along its instruction stream, from `ncu -i X.ncu-rep --page source --csv --print-source sass`.
usage: ncu_stall_hist.py source_page.csv [instructions per bucket]"""
import csv, collections, sys
f = open(sys.argv[1]); f.readline()
rows = list(csv.DictReader(f))
step = int(sys.argv[2]) if len(sys.argv) > 2 else 60
cols = [c for c in rows[0].keys() if c and c.startswith('stall_') and 'Not Issued' not in c]
total = sum(int(r['# Samples']) for r in rows)
tot = collections.Counter()
for r in rows:
    for c in cols:
        try: tot[c[6:]] += int(r[c])
        except ValueError: pass
print("instructions %d, samples %d: %s" % (len(rows), total, ", ".join("%s %.1f%%" % (c, 100.0 * v / total) for c, v in tot.most_common(6))))
for a in range(0, len(rows), step):
    t = collections.Counter(); n = 0; ops = collections.Counter()
    for r in rows[a:a + step]:
        n += int(r['# Samples'])
        for c in cols:
            try: t[c[6:]] += int(r[c])
            except ValueError: pass
        w = r['Source'].split()
        ops[w[1] if w[0].startswith('@') else w[0]] += 1
    if n:
        print("%5d-%5d %5.1f%%  %-58s %s" % (a, a + step, 100.0 * n / total, " ".join("%s %d" % x for x in t.most_common(3) if x[1]), " ".join("%s x%d" % x for x in ops.most_common(4))))
