"""Replay one tools/fuzz_parity.py case and find the first iteration / posterior entry that differs from the oracle."""
import sys, os, subprocess, json
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
case, seed = int(sys.argv[1]), int(sys.argv[2])
import importlib.util
spec = importlib.util.spec_from_file_location("fz", "/root/repo/tools/fuzz_parity.py"); fz = importlib.util.module_from_spec(spec); spec.loader.exec_module(fz)
from nbldpc_b200 import capi
from oracle import pyoracle as po
d = fz.gen_case(case, seed)
code, y, sigma, punct = d["code"], d["y"], d["sigma"], d["punct"]
q, N = code["q"], code["N"]
path = "/tmp/fc.txt"; open(path, "w").write(po.code_text(code))
oc = po.OracleCode(code)
L = np.stack([po.demap_bpsk(q, N, yy, sigma, punct) for yy in y])
for it in range(1, d["max_iter"] + 1):
    want = [oc.decode(d["method"], it, d["p0"], d["p1"], d["factor"], d["offset"], l) for l in L]
    dec = capi.Decoder(capi.Code.from_file(path), d["method"], it, d["p0"], d["p1"], d["factor"], d["offset"], precision=capi.PREC_F64, max_batch=len(L), tie_mode=d["tie"])
    print("paths", dec.info() if hasattr(dec, "info") else "")
    r = dec.decode(L, want_post=True)
    ref = np.stack([w[3] for w in want])
    df = np.abs(r["post"] - ref)
    print("sym equal", np.array_equal(r["sym"], np.stack([w[2] for w in want])), end=" ")
    print("max_iter", it, "max post diff", df.max(), "at", np.unravel_index(df.argmax(), df.shape), "iters", r["iterations"], [w[1] for w in want])
    if df.max() > 1e-9:
        fi = np.unravel_index(df.argmax(), df.shape)[0]
        bad = np.argwhere(df[fi] > 1e-9)
        print("frame", fi, "differing (var, x):", bad[:40].tolist())
        for v, x in bad[:10]:
            print("  var", v, "x", x + 1, "got", repr(r["post"][fi, v, x]), "ref", repr(ref[fi, v, x]))
        os.makedirs("/root/repo/gpurun_out", exist_ok=True)
        np.savez("/root/repo/gpurun_out/fuzz_case_%d_%d.npz" % (seed, case), L=L[fi], got=r["post"][fi], ref=ref[fi], it=it, code=json.dumps(code, default=lambda o: o.tolist() if hasattr(o, "tolist") else int(o)))
        break
