#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source sass` output: stall totals and the
hottest SASS instructions (first kernel in the file).   python tools/ncu_sass_top.py sass.csv [top_n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
his = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
hi = his[0]
end = his[1] - 1 if len(his) > 1 else len(rows)
hdr, data = rows[hi], [r for r in rows[hi + 1:end] if len(r) > 5]
ci = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ci['# Samples']]) for r in data)
print('total samples', tot, 'instrs', len(data), 'warp-instr executed', sum(int(r[ci['Instructions Executed']]) for r in data))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {s: sum(int(r[ci[s]]) for r in data) for s in stalls}
print(sorted(agg.items(), key=lambda x: -x[1])[:8])
idx = sorted(range(len(data)), key=lambda i: -int(data[i][ci['# Samples']]))[:top_n]
for i in sorted(idx):
    r = data[i]
    st = {s: int(r[ci[s]]) for s in stalls if int(r[ci[s]]) > 0}
    top = sorted(st.items(), key=lambda x: -x[1])[:3]
    print(i, r[ci['Source']][:72].ljust(72), r[ci['# Samples']], r[ci['Instructions Executed']], top)
