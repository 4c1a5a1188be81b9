#!/usr/bin/env python
"""stall hot spots from `ncu -i rep --page source --csv --kernel-name ... > file`: usage ncu_hotspots.py file"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr) - 2 and r[ix['# Samples']].isdigit()]
tot = sum(int(r[ix['# Samples']]) for r in data)
print("total samples", tot, " instructions", len(data))
cols = ['stall_long_sb', 'stall_barrier', 'stall_wait', 'stall_math', 'stall_not_selected', 'stall_short_sb', 'stall_lg',
        'stall_mio', 'stall_no_inst', 'stall_dispatch', 'stall_selected', 'stall_branch_resolving']
for col in cols:
    print(f"  {col:24s} {sum(int(r[ix[col]]) for r in data):8d}")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 14
for col in cols[:2]:
    print("--- top", col)
    for r in sorted(data, key=lambda r: -int(r[ix[col]]))[:n]:
        print(f"{r[ix[col]]:>6s} {r[ix['Address']][-5:]} {r[ix['Source']].strip()}")
print("--- top sample lines")
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:2 * n]:
    print(f"{r[ix['# Samples']]:>6s} {r[ix['Address']][-5:]} {r[ix['Source']].strip()}")

# dynamic instruction mix by opcode
import collections
seen, c = set(), collections.Counter()
for r in data:
    a = r[ix['Address']]
    if a in seen:
        continue
    seen.add(a)
    src = r[ix['Source']].strip().split()
    op = (src[1] if src[0].startswith('@') else src[0]).split('.')[0]
    c[op] += int(r[ix['Instructions Executed']])
tot = sum(c.values())
print("--- executed warp instructions", tot)
for k, v in c.most_common(28):
    print(f"{k:10s} {v:12d} {100 * v / tot:5.1f}%")
