#!/usr/bin/env python
"""Static SASS instruction count per source line of one kernel (development aid).

    python tools/sass_by_line.py CUBIN [--file tems_warp.cuh] [--ranges "385-467:VN,474-488:delta"]

Uses `nvdisasm -gi` (needs -lineinfo at compile time).  An instruction is attributed to the OUTERMOST line of its inline chain
that lies in --file (so that helper functions count for the statement that calls them)."""
import argparse, collections, re, subprocess, sys

ap = argparse.ArgumentParser()
ap.add_argument("cubin")
ap.add_argument("--file", default="tems_warp.cuh")
ap.add_argument("--ranges", default="")
ap.add_argument("--ops", action="store_true")
ap.add_argument("--inner", action="store_true")
a = ap.parse_args()
txt = subprocess.run(["nvdisasm", "-gi", "-c", a.cubin], capture_output=True, text=True).stdout
cnt = collections.Counter(); ops = collections.defaultdict(collections.Counter)
inner = collections.defaultdict(collections.Counter)
chain = []; fresh = True
for l in txt.splitlines():
    if "//##" in l:
        if fresh: chain = []; fresh = False
        for m in re.finditer(r'"([^"]+)", line (\d+)', l):
            c = (m.group(1).split("/")[-1], int(m.group(2)))
            if not chain or chain[-1] != c: chain.append(c)
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_]+)', l)
    if m:
        fresh = True
        own = [c for c in chain if c[0] == a.file]
        key = own[-1][1] if own else -1
        cnt[key] += 1; ops[key][m.group(2)] += 1
        if len(own) > 1: inner[key][own[-2][1]] += 1
        elif chain and chain[0][0] != a.file: inner[key][chain[0]] += 1
tot = sum(cnt.values())
print("total", tot)
if a.ranges:
    for r in a.ranges.split(","):
        rg, name = r.split(":"); lo, hi = map(int, rg.split("-"))
        n = sum(v for k, v in cnt.items() if lo <= k <= hi)
        o = collections.Counter()
        for k in cnt:
            if lo <= k <= hi: o.update(ops[k])
        print("%-14s %4d-%4d %5d  %s" % (name, lo, hi, n, dict(o.most_common(8))))
else:
    for k in sorted(cnt):
        print(k, cnt[k], dict(ops[k].most_common(5)) if a.ops else "", dict(inner[k].most_common(12)) if a.inner else "")
