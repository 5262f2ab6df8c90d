#!/usr/bin/env python
"""Sum the warp-stall sampling columns of an `ncu --page source --print-source sass --csv` dump: where do the warps wait?"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
cols = [h for h in hdr if h.startswith('stall_') and '(Not Issued)' not in h]
tot = collections.Counter()
for r in rows[2:]:
    if len(r) < len(hdr): continue
    for c in cols:
        v = r[ix[c]]
        if v.isdigit(): tot[c] += int(v)
s = sum(tot.values())
for c, n in tot.most_common():
    if n: print("%-24s %8d %5.1f%%" % (c, n, 100.0 * n / s))
