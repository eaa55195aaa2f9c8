#!/usr/bin/env python
"""Executed warp instructions per opcode from the SASS page of an ncu report:

    ncu -i X.ncu-rep --page source --csv > src.csv;  python tools/ncu_opcodes.py src.csv [blocks_per_launch]
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
blocks = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
I, SRC = ix['Instructions Executed'], ix['Source']
tot = collections.Counter()
for r in rows[2:]:
    words = r[SRC].strip().split()
    if not words:
        continue
    op = words[1] if words[0].startswith('@') and len(words) > 1 else words[0]
    tot[op.rstrip(';').split('.')[0]] += int(r[I])
total = sum(tot.values())
print('warp instructions per block: %.0f' % (total / blocks))
for k, v in tot.most_common(24):
    print('%-8s %6.2f %%  %8.1f per block' % (k, 100.0 * v / total, v / blocks))
