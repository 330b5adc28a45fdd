"""Summarise an `ncu --page source --csv` export: opcode mix and the SASS lines with most stall samples."""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if r and r[0] == 'Kernel Name':
        break                      # first kernel of the export only
    if len(r) > idx['Instructions Executed'] and r[0] != 'Address':
        data.append(r)
iv = lambda r, k: int(float(r[idx[k]] or 0))
tot = sum(iv(r, 'Instructions Executed') for r in data)
samp = sum(iv(r, '# Samples') for r in data)
print('total warp inst', tot, 'samples', samp, 'sass lines', len(data))
c, s = Counter(), Counter()
for r in data:
    t = r[idx['Source']].split()
    op = t[1] if t[0].startswith('@') else t[0]
    op = op.split('.')[0]
    c[op] += iv(r, 'Instructions Executed')
    s[op] += iv(r, '# Samples')
for op, n in c.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 22):
    print(f"{op:14s} {n:10d} {100*n/tot:5.1f}%  samples {100*s[op]/max(samp,1):5.1f}%")
print('top stall lines')
for r in sorted(data, key=lambda r: -iv(r, '# Samples'))[:24]:
    print(f"{iv(r,'# Samples'):6d} {iv(r,'Instructions Executed'):9d}  {r[idx['Address']][-5:]}  {r[idx['Source']][:100]}")
