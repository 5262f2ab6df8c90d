#!/usr/bin/env python
"""Static SASS opcode histogram of every kernel of libasvgf.so (cuobjdump -sass): which instructions the build contains --
UBLKCP / SYNCS (TMA bulk copy + mbarrier), HMNMX2, MUFU, LDS/STS, LDG widths ...  usage: sass_static_mix.py [libasvgf.so] [filter]"""
import collections, re, subprocess, sys, os
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "raytracedgl1_b200", "libasvgf.so")
flt = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
name, mixes = None, collections.OrderedDict()
for l in out.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0].replace("asvgf::", "").replace("void ", "")
        mixes[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
    if m and name:
        t = m.group(1).split()
        op = t[1] if t[0].startswith("@") else t[0]
        mixes[name][op if op.startswith(("UBLKCP", "SYNCS", "LDG", "STG", "LDS", "STS", "MUFU", "ATOM", "RED", "CCTL")) else op.split(".")[0]] += 1
for name, c in mixes.items():
    if flt and flt not in name:
        continue
    print("%s: %d instructions" % (name, sum(c.values())))
    print("   " + ", ".join("%s %d" % kv for kv in c.most_common(26)))
