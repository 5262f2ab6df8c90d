#!/usr/bin/env python
"""Join an `ncu --page source --print-source sass --csv` dump with nvdisasm line info of the same build:
executed warp-instructions per source line (innermost inlined location).
usage: sass_lines.py <ncu_sass.csv> <cubin> <mangled-kernel-substring> [top]"""
import csv, sys, subprocess, collections, re
ncu_csv, cubin, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
lines = []; cur = None; on = False
for l in dis:
    if l.startswith("//---") and ".text." in l:
        on = kern in l
        continue
    if not on: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l): lines.append(cur)
rows = list(csv.reader(open(ncu_csv)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
ex = []
for r in rows[2:]:
    if len(r) < len(hdr) or not r[ix["Instructions Executed"]].isdigit(): 
        if ex: break
        continue
    ex.append((int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]]), r[ix["Source"]].strip()))
print("sass instrs: ncu %d nvdisasm %d" % (len(ex), len(lines)))
agg = collections.Counter(); smp = collections.Counter()
for (n, s, _), loc in zip(ex, lines):
    agg[loc] += n; smp[loc] += s
tot = sum(agg.values()); ts = sum(smp.values())
print("total warp instr %d, samples %d" % (tot, ts))
src = {}
for loc, n in agg.most_common(top):
    f, ln = loc
    if f not in src:
        try: src[f] = open("/root/repo/raytracedgl1_b200/csrc/" + f).read().splitlines()
        except Exception: src[f] = []
    text = src[f][ln - 1].strip()[:110] if ln - 1 < len(src[f]) else ""
    print("%5.1f%% %5.1f%%s %s:%d  %s" % (100.0 * n / tot, 100.0 * smp[loc] / max(ts, 1), f, ln, text))
