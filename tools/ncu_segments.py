#!/usr/bin/env python
"""Group the SASS of an ncu source page into runs of instructions with the same execution count: shows which code
path costs how many warp-instructions and how many lanes are active there.  usage: ncu_segments.py src.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
isrc, iss, iex, it = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed')
data = []
for r in rows[2:]:
    try:
        data.append((int(r[iex]), int(r[it]), int(r[iss]), r[isrc].strip()[:70]))
    except ValueError:
        continue
tw = sum(d[0] for d in data)
print('total warp-inst %.2fe9 thread-inst %.2fe9 n=%d' % (tw / 1e9, sum(d[1] for d in data) / 1e9, len(data)))
seg, cur = [], None
for i, (w, t, s, src) in enumerate(data):
    if w < float(sys.argv[2]) if len(sys.argv) > 2 else w < 2e6:
        if cur:
            seg.append(cur)
            cur = None
        continue
    if cur and abs(cur['w'] - w) < 0.03 * w and abs(cur['t'] - t) < 0.03 * max(t, 1):
        cur['n'] += 1
        cur['smp'] += s
        cur['end'] = i
    else:
        if cur:
            seg.append(cur)
        cur = {'start': i, 'end': i, 'w': w, 't': t, 'n': 1, 'smp': s, 'first': src}
if cur:
    seg.append(cur)
for sg in seg:
    print('%4d-%4d n=%3d  warp %6.1fM thr/warp %4.1f  tot warp-inst %6.2fe9 smp %7d  %s' % (sg['start'], sg['end'], sg['n'], sg['w'] / 1e6, sg['t'] / max(sg['w'], 1), sg['n'] * sg['w'] / 1e9, sg['smp'], sg['first']))
