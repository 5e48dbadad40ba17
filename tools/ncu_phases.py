#!/usr/bin/env python
"""Per-phase profile of the T-EMS warp kernel from an ncu source page (development aid): dynamic instructions per check row, static
size, share of the warp stall samples and the leading stall reasons.  Phases are found from the section comments of tems_warp.cuh.

    python tools/ncu_phases.py SOURCE.csv CUBIN --rows ROWS"""
import argparse, collections, csv, os, re, subprocess
ap = argparse.ArgumentParser()
ap.add_argument("source_csv"); ap.add_argument("cubin"); ap.add_argument("--rows", type=float, required=True)
ap.add_argument("--file", default="tems_warp.cuh")
ap.add_argument("--kernel", default="", help="substring of the mangled kernel name: take only that .text section of a multi-kernel cubin")
a = ap.parse_args()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(ROOT, "nbldpc_b200", "csrc", a.file)).read().splitlines()
marks = [("setup", "__global__ void"), ("VN", "steps A-C for the row"), ("delta", "delta domain (TEMS"), ("colorder", "column order per eta"),
         ("search", "pair search: deltaW"), ("path", "the recorded path of each owned"), ("trellis", "output trellis Tout"),
         ("bstems", "BS-TEMS (Decoding_BS"), ("fillin", "fill-in (NBLDPC"), ("output", "normal domain, un-permute")]
starts = []
for name, pat in marks:
    for i, l in enumerate(src):
        if pat in l and (name != "setup" or "k_iterate_tems_warp" in src[i] or "k_iterate_tems_warp" in src[min(i + 1, len(src) - 1)]):
            starts.append((i + 1, name)); break
starts.sort()
def phase(line):
    cur = "helpers"
    for ln, name in starts:
        if line >= ln: cur = name
    return cur
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
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+', l):
        fresh = True
        own = [c[1] for c in chain if c[0] == a.file]
        keys.append(phase(own[-1]) if own else "other")
rows = list(csv.reader(open(a.source_csv)))
hdr = rows[1]; data = rows[2:]; col = {n: i for i, n in enumerate(hdr)}
assert len(keys) == len(data), (len(keys), len(data))
st = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot = sum(float(r[col["# Samples"]]) for r in data)
order = ["setup", "VN", "delta", "colorder", "search", "path", "trellis", "fillin", "output", "other", "helpers", "bstems"]
agg = {k: [0, 0.0, 0.0, collections.Counter()] for k in order}
for r, k in zip(data, keys):
    v = agg[k]; v[0] += 1; v[1] += float(r[col["Instructions Executed"]]); v[2] += float(r[col["L1 Wavefronts Shared"]])
    for x in st: v[3][x] += float(r[col[x]])
print("%-9s %6s %8s %8s %8s  leading stall reasons (%% of the phase's samples)" % ("phase", "static", "instr/row", "smem wf", "samples"))
T = [0, 0.0, 0.0, collections.Counter()]
for k in order:
    v = agg[k]
    if not v[0]: continue
    t = sum(v[3].values())
    T[0] += v[0]; T[1] += v[1]; T[2] += v[2]; T[3].update(v[3])
    print("%-9s %6d %8.0f %8.0f %7.1f%%  %s" % (k, v[0], v[1] / a.rows, v[2] / a.rows, 100 * t / tot, " ".join("%s %.0f" % (n[6:], 100 * c / max(t, 1)) for n, c in v[3].most_common(5))))
t = sum(T[3].values())
print("%-9s %6d %8.0f %8.0f %7.1f%%  %s" % ("all", T[0], T[1] / a.rows, T[2] / a.rows, 100 * t / tot, " ".join("%s %.0f" % (n[6:], 100 * c / t) for n, c in T[3].most_common(7))))
