"""latency of nbl_decode_batch for small batches (development aid): the reference's per-frame call is B = 1"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
from nbldpc_b200 import capi
from config_sweep import write_code_files
os.makedirs("/tmp/nbldpc_sweep", exist_ok=True)
paths, cons = write_code_files("/tmp/nbldpc_sweep")
code = capi.Code.from_file(paths["divsalar.UNBLDPC.512.256.GF.256"])
rng = np.random.default_rng(1)
for B in (1, 8, 148, 1184):
    for es in (1, 0):
        dec = capi.Decoder(code, capi.METHOD_TEMS, 10, 2, 2, 1.0, 0.5, precision=capi.PREC_F32, early_stop=es, max_batch=B)
        y = 1.0 + 0.75 * rng.standard_normal((B, 512))
        lch = dec.demap_bpsk(y, 0.75)
        for _ in range(3): dec.decode(lch)
        t0 = time.perf_counter()
        n = 20
        for _ in range(n): r = dec.decode(lch)
        dt = (time.perf_counter() - t0) / n
        print("B %5d early_stop %d: %8.1f us per call (host wall clock, host buffers), %.2f us per frame, mean iterations %.1f" % (B, es, dt * 1e6, dt * 1e6 / B, r["iterations"].mean()))
