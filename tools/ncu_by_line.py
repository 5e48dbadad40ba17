#!/usr/bin/env python
"""Dynamic per-source-line profile of one kernel: joins the SASS page of an ncu report (`ncu -i X.ncu-rep --page source --csv`)
with the line table of the same kernel's cubin (`nvdisasm -gi`), by instruction index.  Development aid.

    python tools/ncu_by_line.py SOURCE.csv CUBIN --rows ROWS [--file tems_warp.cuh] [--ranges "458-465:VN,..."]

Prints, per source line (outermost line of the inline chain inside --file) or per named range: warp instructions executed
per check row, shared-memory wavefronts per row (and the ideal count), stall samples."""
import argparse, collections, csv, re, subprocess

ap = argparse.ArgumentParser()
ap.add_argument("source_csv"); ap.add_argument("cubin")
ap.add_argument("--rows", type=float, required=True, help="check rows the launch processed (frames x M)")
ap.add_argument("--file", default="tems_warp.cuh")
ap.add_argument("--ranges", default="")
ap.add_argument("--inner", action="store_true", help="break each line down by the innermost line of the chain")
ap.add_argument("--opmix", default="", help="LO-HI: dynamic opcode mix (per row) of the instructions of these lines")
ap.add_argument("--kernel", default="", help="substring of the mangled kernel name: take only that .text section of a multi-kernel cubin")
a = ap.parse_args()

txt = subprocess.run(["nvdisasm", "-gi", "-c", a.cubin], capture_output=True, text=True).stdout.splitlines()
if a.kernel:
    sel = []; on = False
    for l in txt:
        if re.match(r"\s*\.section\s+\.text\.", l): on = a.kernel in l
        elif re.match(r"\s*\.section\s", l): on = False
        if on: sel.append(l)
    txt = sel
keys = []; chain = []; fresh = True
for l in txt:
    if "//##" in l:
        if fresh: chain = []; fresh = False
        for m in re.finditer(r'"([^"]+)", line (\d+)', l):
            c = (m.group(1).split("/")[-1], int(m.group(2)))
            if not chain or chain[-1] != c: chain.append(c)
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_]+)', l)
    if m:
        fresh = True
        own = [c[1] for c in chain if c[0] == a.file]
        keys.append((own[-1] if own else -1, own[0] if own else -1, m.group(2)))
rows = list(csv.reader(open(a.source_csv)))
hdr = rows[1]; data = rows[2:]
col = {n: i for i, n in enumerate(hdr)}
assert len(data) == len(keys), (len(data), len(keys))
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0.0])
inner = collections.defaultdict(lambda: collections.defaultdict(float))
for r, (outer, inn, op) in zip(data, keys):
    v = agg[outer]
    ie = float(r[col["Instructions Executed"]])
    v[0] += ie; v[1] += float(r[col["L1 Wavefronts Shared"]]); v[2] += float(r[col["L1 Wavefronts Shared Ideal"]])
    v[3] += float(r[col["# Samples"]])
    inner[outer][inn] += ie
if a.opmix:
    lo, hi = map(int, a.opmix.split("-"))
    mix = collections.Counter()
    for r, (outer, inn, op) in zip(data, keys):
        if lo <= outer <= hi: mix[op] += float(r[col["Instructions Executed"]]) / a.rows
    print("opcode mix of lines %s: total %.1f per row" % (a.opmix, sum(mix.values())))
    for k, v in mix.most_common(): print("   %-12s %8.1f" % (k, v))
tot = [sum(v[i] for v in agg.values()) for i in range(4)]
print("per check row: %.0f instructions, %.0f smem wavefronts (%.0f ideal); %d samples" % (tot[0] / a.rows, tot[1] / a.rows, tot[2] / a.rows, tot[3]))
def show(name, v, extra=""):
    print("%-16s instr %7.1f  wavefronts %7.1f (ideal %6.1f)  samples %5.1f%% %s" % (name, v[0] / a.rows, v[1] / a.rows, v[2] / a.rows, 100 * v[3] / tot[3], extra))
if a.ranges:
    for rg in a.ranges.split(","):
        r, name = rg.split(":"); lo, hi = map(int, r.split("-"))
        v = [sum(x[i] for k, x in agg.items() if lo <= k <= hi) for i in range(4)]
        show(name, v)
else:
    for k in sorted(agg):
        ex = ""
        if a.inner:
            ex = " ".join("%d:%.0f" % (i, n / a.rows) for i, n in sorted(inner[k].items(), key=lambda kv: -kv[1])[:8] if i != k)
        show(str(k), agg[k], ex)
