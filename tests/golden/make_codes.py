"""Regenerate tests/golden/codes.json from the reference's data files (run in the build container only).

The reference ships its parity-check matrices, the H_bit images and the QAM constellations as text files
that cannot be regenerated.  /root/reference does not exist on the GPU box, so the
tests, smoke() and bench.py read this bundle instead and re-emit the reference text formats through
oracle.pyoracle.code_text() when they exercise the product's file loaders.
"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle.pyoracle import parse_code_text  # noqa: E402

REF = os.environ.get("NBLDPC_REFERENCE", "/root/reference")
CODES = ["divsalar.UNBLDPC.128.64.GF.16", "divsalar.UNBLDPC.256.128.GF.16", "divsalar.UNBLDPC.512.256.GF.16",
         "divsalar.UNBLDPC.128.64.GF.256", "divsalar.UNBLDPC.256.128.GF.256", "divsalar.UNBLDPC.512.256.GF.256",
         "divsalar.CNBLDPC.128.64.GF.256", "divsalar.CNBLDPC.256.128.GF.256", "divsalar.CNBLDPC.512.256.GF.256",
         "BDS.576.288.GF.64"]


def main():
    bundle = {"codes": {}, "constellations": {}, "h_bit": {}}
    for name in CODES:
        c = parse_code_text(open(os.path.join(REF, name + ".txt")).read())
        bundle["codes"][name] = {"N": c["N"], "M": c["M"], "q": c["q"],
                                 "var": [[list(e) for e in l] for l in c["var"]],
                                 "chk": [[list(e) for e in l] for l in c["chk"]]}
    for name in ["BPSK", "GRAY_64QAM", "GRAY_256QAM"]:
        tok = open(os.path.join(REF, name + ".txt")).read().split()
        pts = {}
        for i in range(0, len(tok), 6):          # "Point: idx Real: re Imag: im" (Comm.cpp:117-126)
            pts[int(tok[i + 1])] = [float(tok[i + 3]), float(tok[i + 5])]
        bundle["constellations"][name] = [pts[i] for i in range(len(pts))]
    rows = [l.split() for l in open(os.path.join(REF, "H_bit_GF256.txt")).read().splitlines() if l.strip()]
    bundle["h_bit"]["H_bit_GF256"] = ["".join(r) for r in rows]
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "codes.json")
    json.dump(bundle, open(out, "w"), separators=(",", ":"))
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
