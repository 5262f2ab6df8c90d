#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source sass --csv` dump by opcode: executed warp instructions and stall samples."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); samples = collections.Counter(); tot = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[ix["Source"]].strip()
    toks = src.split()
    if not toks: continue
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0] if len(sys.argv) < 3 else op
    if not r[ix["Instructions Executed"]].isdigit(): continue
    n = int(r[ix["Instructions Executed"]]); s = int(r[ix["# Samples"]])
    ops[op] += n; samples[op] += s; tot += n
print("total warp instr", tot, "static", len(rows) - 2)
for op, n in ops.most_common(28):
    print("%-14s %12d %5.1f%%  samples %d" % (op, n, 100.0 * n / tot, samples[op]))
