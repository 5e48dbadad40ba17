import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from nbldpc_b200 import capi
from oracle import pyoracle as po
from test_warp_paths_gpu import _random_code
import tempfile, pathlib
rng = np.random.default_rng(11)
tmp = pathlib.Path(tempfile.mkdtemp())
for q, regular, method, p0, p1 in [(64, True, 4, 2, 2), (32, False, 4, 2, 2), (256, True, 4, 2, 2), (16, False, 4, 2, 2), (64, True, 2, 8, 3), (16, False, 1, 0, 0)]:
    code = _random_code(q, 8, rng, regular)
    path = tmp / ("c%d.txt" % q); path.write_text(po.code_text(code))
    oc = po.OracleCode(code)
    F = 12
    lch = rng.integers(-6, 3, size=(F, code["N"], q - 1)).astype(np.float64)     # integer LLRs: exact ties everywhere
    lch[:, ::3, :] = 0.0                                                      # every third symbol "punctured"
    want = [oc.decode(method, 5, p0, p1, 1.0, 0.5 if method != 1 else 0.0, l) for l in lch]
    dec = capi.Decoder(capi.Code.from_file(str(path)), method, 5, p0, p1, 1.0, 0.5 if method != 1 else 0.0, precision=capi.PREC_F64, max_batch=F)
    r = dec.decode(lch, want_post=True)
    okc = np.array_equal(r["converged"], np.array([w[0] for w in want], np.int32))
    oki = np.array_equal(r["iterations"], np.array([w[1] for w in want], np.int32))
    oks = (r["sym"] == np.stack([w[2] for w in want])).all(axis=1)
    ref = np.stack([w[3] for w in want])
    err = np.abs(r["post"] - ref).max(axis=(1, 2))
    print("q", q, "method", method, "conv", okc, "iters", oki, "sym frames ok", oks.sum(), "/", F, "post err", err.max())
