#!/usr/bin/env python
"""Per-phase (between CTA barriers) share of warp-stall samples and executed instructions of the first kernel in an
`ncu --page source --csv` dump (optionally gzipped): scripts/ncu_phases.py file.csv[.gz]"""
import csv, gzip, io, sys
fn = sys.argv[1]
f = io.TextIOWrapper(gzip.open(fn)) if fn.endswith(".gz") else open(fn)
rows = list(csv.reader(f))
hdr = rows[1]
iS, iSrc, iEx = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
body = []
for r in rows[2:]:
    if len(r) < len(hdr):
        break
    body.append(r)
tot = sum(int(r[iS]) for r in body); totx = sum(int(r[iEx]) for r in body)
print("SASS instructions", len(body), "samples", tot, "warp instructions executed", totx)
stalls = {h: hdr.index(h) for h in hdr if h.startswith('stall_') and 'Not Issued' not in h}
acc_s = acc_x = 0; start = 0; agg = {}
def flush(i):
    global acc_s, acc_x, start, agg
    if acc_s or acc_x:
        t = sorted(agg.items(), key=lambda kv: -kv[1])[:4]
        print(f"[{start:5d}-{i:5d}] samples {acc_s/tot*100:5.1f}%  instr {acc_x/totx*100:5.1f}% ",
              [(k.replace('stall_', ''), round(v / max(acc_s, 1) * 100)) for k, v in t])
    acc_s = acc_x = 0; start = i + 1; agg = {}
for i, r in enumerate(body):
    acc_s += int(r[iS]); acc_x += int(r[iEx])
    for h, j in stalls.items():
        v = int(r[j])
        if v:
            agg[h] = agg.get(h, 0) + v
    if 'BAR.SYNC' in r[iSrc]:
        flush(i)
flush(len(body) - 1)
