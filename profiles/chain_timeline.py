#!/usr/bin/env python
"""Per-instruction cycle attribution of a latency-bound kernel from an ncu source-page CSV
(ncu -i X.ncu-rep --page source --csv --print-source sass): with ONE warp per SM sub-partition every cycle of the warp is a
sample in exactly one state, so samples / total x cycles-per-step is the time the chain spends at each instruction.
usage: chain_timeline.py src.csv <units = warps x steps> <cycles per step> [out.txt]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
units, cps = float(sys.argv[2]), float(sys.argv[3])
tot = sum(int(r[ix['# Samples']]) for r in data)
agg = {s: sum(int(r[ix[s]]) for r in data) for s in stalls}
print('samples %d; by state: %s' % (tot, ', '.join('%s %.1f%%' % (s.replace('stall_', ''), 100 * v / tot)
                                                    for s, v in sorted(agg.items(), key=lambda x: -x[1])[:7])))
out, cum = [], 0.
for r in data:
    ex = int(r[ix['Instructions Executed']])
    if ex < units * 0.1:
        continue
    smp = int(r[ix['# Samples']])
    cyc = smp / tot * cps
    cum += cyc
    top = max(stalls, key=lambda s: int(r[ix[s]]))
    out.append('%s c=%7.1f +%5.1f x%.2f %-10s %s' % (r[0][-5:], cum, cyc, ex / units, top.replace('stall_', ''), r[1].strip()[:64]))
print('instructions in the loop: %d, cycles attributed: %.0f of %.0f' % (len(out), cum, cps))
if len(sys.argv) > 4:
    open(sys.argv[4], 'w').write('\n'.join(out) + '\n')
