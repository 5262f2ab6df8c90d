#!/usr/bin/env python
"""Aggregate an `ncu -i X.ncu-rep --page source --print-source sass --csv` dump by opcode: executed warp instructions and stall
samples, per kernel result in the dump (a dump may hold several launches; each starts with a "Kernel Name" row, and each
instruction is listed once -- the dump repeats the listing, the repetition is dropped).  usage: sass_mix.py <dump.csv> [full-opcode]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
full = len(sys.argv) > 2
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and len(r) >= len(cur["hdr"]):
        cur["rows"].append(r)
prev = None
for b in blocks:
    key = (b["name"], tuple(tuple(r[:8]) for r in b["rows"][:64]))
    if key == prev:          # the dump lists every result twice (SASS view and its repetition)
        prev = None
        continue
    prev = key
    ix = {h: i for i, h in enumerate(b["hdr"])}
    ops, samples, tot, seen = collections.Counter(), collections.Counter(), 0, set()
    for r in b["rows"]:
        if r[ix["Address"]] in seen:
            continue
        seen.add(r[ix["Address"]])
        toks = r[ix["Source"]].strip().split()
        if not toks or not r[ix["Instructions Executed"]].isdigit():
            continue
        op = toks[1] if toks[0].startswith("@") else toks[0]
        op = op if full else op.split(".")[0]
        n, s = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
        ops[op] += n; samples[op] += s; tot += n
    if not tot:
        continue
    print("== %s: %d executed warp instructions, %d static" % (b["name"].replace("asvgf::", ""), tot, len(seen)))
    for op, n in ops.most_common(22):
        print("   %-12s %12d %5.1f%%  stall samples %d" % (op, n, 100.0 * n / tot, samples[op]))
