"""Stall samples per CUDA source line of one kernel, from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.
usage: ncu_line_hist.py correlated.csv kernel-substring file-substring [top N]"""
import csv, sys
lines = open(sys.argv[1]).read().split('\n')
kern, fsub = sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
blocks = [i for i, l in enumerate(lines) if l.startswith('"File Path"')]
acc = {}
for bi, b in enumerate(blocks):
    if fsub not in lines[b] or kern not in lines[b + 1]:
        continue
    end = blocks[bi + 1] if bi + 1 < len(blocks) else len(lines)
    hdr = next(csv.reader([lines[b + 2]]))
    si = hdr.index('# Samples')
    for r in csv.reader(lines[b + 3:end]):
        if len(r) > si and r[0] != '':
            try:
                acc[(int(r[0]), r[1])] = acc.get((int(r[0]), r[1]), 0) + int(r[si])
            except ValueError:
                pass
tot = sum(acc.values())
print("samples in %s (%s): %d" % (kern, fsub, tot))
for (ln, src), n in sorted(acc.items(), key=lambda kv: -kv[1])[:top]:
    print("%5.1f%% %5d  %s" % (100.0 * n / max(tot, 1), ln, src[:130]))
