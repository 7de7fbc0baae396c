#!/usr/bin/env python
"""Aggregate ncu warp-stall samples by CUDA source line:
   ncu -i X.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:K > src.csv
   python profiles/src_hot.py src.csv <steps-for-normalisation> [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
norm = float(sys.argv[2]) if len(sys.argv) > 2 else 1.
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
cur, agg, tot, totex = None, {}, 0, 0
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if len(r) < 8 or not r[0].strip().isdigit():
        continue
    try:
        s, ex = int(r[6]), int(r[7])
    except ValueError:
        continue
    a = agg.setdefault((cur, int(r[0]), r[1].strip()[:100]), [0, 0])
    a[0] += s; a[1] += ex; tot += s; totex += ex
print('total samples %d, instructions/unit %.1f' % (tot, totex / norm))
for k, (s, ex) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print('%5.1f%%  %7.1f i/unit  %s:%d  %s' % (100. * s / max(tot, 1), ex / norm, k[0], k[1], k[2]))
