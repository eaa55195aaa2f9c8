#!/usr/bin/env python
"""Executed warp instructions per opcode from the SASS page of an ncu report:

    ncu -i X.ncu-rep --page source --csv > src.csv;  python tools/ncu_opcodes.py src.csv [blocks_per_launch] [out.json]

With a third argument the counts also go to a JSON file, together with the EXECUTED floating-point operations per block
(32 lanes x (4 per FFMA2, 2 per FADD2 / FMUL2 / FFMA / DFMA, 1 per FADD / FMUL / DADD / DMUL)) that bench.py quotes beside
the algorithmic 5 N log2 N count."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
blocks = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
hdr = rows[1] if 'Source' in rows[1] else rows[0]
rows = rows if 'Source' in rows[1] else [None] + rows
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
if len(sys.argv) > 3:
    import json
    flop = {'FFMA2': 4, 'FADD2': 2, 'FMUL2': 2, 'FFMA': 2, 'DFMA': 2, 'FADD': 1, 'FMUL': 1, 'DADD': 1, 'DMUL': 1}
    per_block = {k: v / blocks for k, v in tot.items()}
    flops = 32 * sum(per_block.get(k, 0.0) * f for k, f in flop.items())
    json.dump({'source': sys.argv[1], 'blocks_per_launch': blocks, 'warp_instructions_per_block': total / blocks,
               'executed_flops_per_block': flops, 'flops_per_opcode': flop,
               'warp_instructions_per_block_by_opcode': {k: round(v, 1) for k, v in sorted(per_block.items(), key=lambda kv: -kv[1])[:40]}},
              open(sys.argv[3], 'w'), indent=1)
    print('executed flops per block: %.0f' % flops)
