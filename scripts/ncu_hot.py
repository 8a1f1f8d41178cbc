#!/usr/bin/env python
"""Top stall-sample instructions of the first kernel in an `ncu --page source --csv` dump."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
iS, iSrc, iEx = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
body = []
for r in rows[2:]:
    if len(r) < len(hdr):
        break  # next kernel
    body.append(r)
tot = sum(int(r[iS]) for r in body)
print('total samples', tot, 'instructions', len(body))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(int(r[hdr.index(h)]) for r in body) for h in stalls}
print('stall totals:', sorted(agg.items(), key=lambda kv: -kv[1])[:8])
idx = sorted(range(len(body)), key=lambda i: -int(body[i][iS]))[:top]
for i in sorted(idx):
    r = body[i]
    st = {h: int(r[hdr.index(h)]) for h in stalls if int(r[hdr.index(h)]) > 0}
    t3 = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(i, r[iSrc].strip()[:70], r[iS], r[iEx], t3)
