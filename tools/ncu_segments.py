#!/usr/bin/env python
"""Split the SASS page of an ncu report (ncu -i X.ncu-rep --page source --csv) at barrier-like
instructions and print, per code segment, its share of stall samples and executed instructions."""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
S, I, SRC = ix['# Samples'], ix['Instructions Executed'], ix['Source']
tot = sum(int(r[S]) for r in data)
toti = sum(int(r[I]) for r in data)
print('total samples', tot, 'warp instr', toti, 'sass rows', len(data))
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]


def new(n):
    return dict(start=n, samples=0, instr=0, st={h: 0 for h in stall_cols}, ops={})


seg, cur = [], new(0)
for n, r in enumerate(data):
    op = r[SRC].strip()
    cur['samples'] += int(r[S])
    cur['instr'] += int(r[I])
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_]+)', op)
    if m:
        cur['ops'][m.group(2)] = cur['ops'].get(m.group(2), 0) + int(r[I])
    for h in stall_cols:
        cur['st'][h] += int(r[ix[h]])
    if re.search(r'\bBAR\.|WARPSYNC|SYNCS|EXIT|UBLKCP', op):
        cur['end'], cur['endn'] = op, n
        seg.append(cur)
        cur = new(n + 1)
seg.append(cur)
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
for s in seg:
    if s['samples'] < tot * thr / 100:
        continue
    top = sorted(s['st'].items(), key=lambda kv: -kv[1])[:4]
    ops = sorted(s['ops'].items(), key=lambda kv: -kv[1])[:5]
    print('%5d-%5d samp %5.1f%% instr %5.1f%%  end=%-34s %s | %s' % (
        s['start'], s.get('endn', -1), 100 * s['samples'] / tot, 100 * s['instr'] / toti, s.get('end', '')[:34],
        ' '.join('%s=%d' % (k[6:], v) for k, v in top), ' '.join('%s=%d' % (k, v // 1000) for k, v in ops)))
