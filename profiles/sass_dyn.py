#!/usr/bin/env python
"""Join the static scheduling info of a kernel's SASS (cuobjdump -sass of the .so) with the per-instruction execution
counts / stall samples of an ncu source-page CSV (--print-source cuda,sass --csv) of the SAME build.
usage: sass_dyn.py lib.sass <mangled-substring> src.csv <units> [list]"""
import re, csv, sys
from collections import Counter
txt = open(sys.argv[1]).read()
part = [p for p in txt.split('Function : ')[1:] if sys.argv[2] in p.split('\n')[0]][0]
pat = re.compile(r'^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/')
lines = part.split('\n')
ins = []
i = 0
while i < len(lines):
    m = pat.match(lines[i])
    if m and i + 1 < len(lines):
        m2 = re.search(r'/\* (0x[0-9a-f]+) \*/', lines[i + 1])
        if m2:
            hi = int(m2.group(1), 16)
            ins.append((int(m.group(1), 16), m.group(2).strip(), (hi >> 41) & 0xF, (hi >> 46) & 7, (hi >> 49) & 7, (hi >> 52) & 0x3F))
            i += 2
            continue
    i += 1
rows = list(csv.reader(open(sys.argv[3])))
units = float(sys.argv[4])
seen, sass = set(), []
for r in rows:
    if len(r) >= 8 and r[0].strip() == '' and r[2].startswith('0x') and r[2] not in seen:
        seen.add(r[2])
        try:
            sass.append((int(r[2], 16), r[3].strip(), int(r[6]), int(r[7])))
        except ValueError:
            pass
sass.sort()
n = min(len(ins), len(sass))
mism = sum(1 for k in range(n) if ins[k][1].split()[0] != sass[k][1].split()[0])
print('static instrs %d / ncu rows %d / opcode mismatches %d' % (len(ins), len(sass), mism))
dyn = sum(s[3] for s in sass[:n]) / units
stall = sum(ins[k][2] * sass[k][3] for k in range(n)) / units
samples = sum(s[2] for s in sass[:n])
print('dynamic instrs/unit %.1f ; sum(static stall x exec)/unit = %.0f cycles ; samples %d' % (dyn, stall, samples))
if len(sys.argv) > 5:
    for k in range(n):
        if sass[k][3] >= units * 0.5:
            a, s = ins[k], sass[k]
            print('%05x st=%2d wb=%d rb=%d wait=%02x ex=%5.1f smp=%5d  %s' % (a[0], a[2], a[3], a[4], a[5], s[3] / units, s[2], a[1][:80]))
