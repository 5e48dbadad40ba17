#!/usr/bin/env python
"""Development probe: the one EMS case of tools/fuzz_parity.py that differs from the oracle (seed 51, case 75; DESIGN.md 2, "EMS and the
same residue") through both kernels and both precisions -- they agree with each other, not with the reference's running sums."""
import sys, os, json
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import importlib.util
spec = importlib.util.spec_from_file_location("fz", "/root/repo/tools/fuzz_parity.py"); fz = importlib.util.module_from_spec(spec); spec.loader.exec_module(fz)
from nbldpc_b200 import capi
from oracle import pyoracle as po
d = fz.gen_case(75, 51)
code, y, sigma, punct = d["code"], d["y"], d["sigma"], d["punct"]
q, N = code["q"], code["N"]
print("chk degrees", [len(l) for l in code["chk"]], "var degrees", [len(l) for l in code["var"]])
path = "/tmp/fc.txt"; open(path, "w").write(po.code_text(code))
oc = po.OracleCode(code)
L = np.stack([po.demap_bpsk(q, N, yy, sigma, punct) for yy in y])
it = 5
want = [oc.decode(d["method"], it, d["p0"], d["p1"], d["factor"], d["offset"], l) for l in L]
ref = np.stack([w[3] for w in want])
for kern in (0, 1):
    for prec in (capi.PREC_F64, capi.PREC_F32):
        dec = capi.Decoder(capi.Code.from_file(path), d["method"], it, d["p0"], d["p1"], d["factor"], d["offset"], precision=prec, max_batch=len(L), kernel=kern)
        r = dec.decode(L, want_post=True)
        df = np.abs(r["post"] - ref)
        print("kernel", dec.kernel(), "prec", prec, "max diff", df.max(), "sym equal", np.array_equal(r["sym"], np.stack([w[2] for w in want])), "iters", r["iterations"])
# the frame alone, and per max_iter with group kernel
