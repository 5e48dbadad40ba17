#!/usr/bin/env python
"""SASS opcode histogram of the shipped iteration kernels, from the in-tree library (evidence for DESIGN.md: which instructions the
hot kernels are made of -- FADD2 / FMNMX3 / REDUX, no tensor or TMA instructions by design).

    python tools/sass_histogram.py > profiles/NAME_sass_opcodes.txt"""
import collections, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "nbldpc_b200", "libnbldpc_b200.so")
KERNELS = [("T-EMS GF(256), one warp per check row (headline, cfg3 / cfg5)", "k_iterate_tems_warpIfLi256ELi4ELi2ELi0"),
           ("QSPA GF(256), one warp per check row (cfg4)", "k_iterate_qspa_warpIfLi256ELi4ELi2"),
           ("EMS GF(64), one warp per check row (cfg2)", "k_iterate_ems_warpIfLi64ELi4ELi2"),
           ("QSPA GF(16), four lanes per check row (cfg1)", "k_iterate_qspa_halfIf")]
with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "decoder.sm_100a.cubin", LIB], cwd=tmp, check=True, capture_output=True)
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(tmp, "decoder.sm_100a.cubin")], capture_output=True, text=True).stdout
cur = None; hist = collections.defaultdict(collections.Counter); arch = set()
for l in sass.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = m.group(1); continue
    m = re.search(r"arch = (\S+)", l)
    if m: arch.add(m.group(1))
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", l)
    if m and cur: hist[cur][m.group(1)] += 1
print("cuobjdump -sass of nbldpc_b200/libnbldpc_b200.so (arch %s); static instruction counts per opcode" % ", ".join(sorted(arch)))
for title, pat in KERNELS:
    for fn, h in hist.items():
        if pat in fn:
            tot = sum(h.values())
            print("\n%s\n  %s: %d instructions (%.1f KB)" % (title, fn, tot, tot * 16 / 1024))
            print("  " + "  ".join("%s %d" % kv for kv in h.most_common()))
            print("  tensor / TMA / TMEM instructions (UTC*MMA, HMMA, UTMALDG, LDTM): %d -- none by design: no dense contraction on this path" %
                  sum(v for k, v in h.items() if k.startswith(("UTC", "HMMA", "UTMA", "LDTM", "STTM", "UBLKCP"))))
