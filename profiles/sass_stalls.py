#!/usr/bin/env python
"""Decode the static scheduling information (stall count, scoreboard barriers) of a kernel's SASS
(cuobjdump -sass output) and print the longest backward-branch loop with per-instruction stall counts.
Volta+ control bits live in the upper 64-bit word: stall = bits 41-44, yield 45, wbar 46-48, rbar 49-51,
wait mask 52-57."""
import re, sys
lines = open(sys.argv[1]).read().split('\n')
ins = []
i = 0
pat = re.compile(r'^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/')
while i < len(lines):
    m = pat.match(lines[i])
    if m and i + 1 < len(lines):
        m2 = re.search(r'/\* (0x[0-9a-f]+) \*/', lines[i + 1])
        if m2:
            hi = int(m2.group(1), 16)
            ins.append(dict(addr=int(m.group(1), 16), text=m.group(2).strip(), stall=(hi >> 41) & 0xF, yld=(hi >> 45) & 1,
                            wbar=(hi >> 46) & 7, rbar=(hi >> 49) & 7, wait=(hi >> 52) & 0x3F))
            i += 2
            continue
    i += 1
# find backward branches
best = None
for k, a in enumerate(ins):
    m = re.search(r'BRA\S*\s+(?:\S+,\s*)?(0x[0-9a-f]+)', a['text'])
    if m:
        tgt = int(m.group(1), 16)
        if tgt < a['addr']:
            span = a['addr'] - tgt
            if best is None or span > best[0]:
                best = (span, tgt, a['addr'])
print('instructions', len(ins), 'loop', [hex(x) for x in best[1:]] if best else None)
lo, hi_ = best[1], best[2]
body = [a for a in ins if lo <= a['addr'] <= hi_]
tot = sum(a['stall'] for a in body)
print('loop body instrs %d, sum of static stall counts %d cycles' % (len(body), tot))
if len(sys.argv) > 2:
    for a in body:
        print('%05x st=%2d y=%d w=%d r=%d wait=%02x  %s' % (a['addr'], a['stall'], a['yld'], a['wbar'], a['rbar'], a['wait'], a['text'][:90]))
