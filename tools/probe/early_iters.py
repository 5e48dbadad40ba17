"""per-iteration kernel times of the headline configuration with the reference's early stop (development aid)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from nbldpc_b200 import capi
from config_sweep import write_code_files
os.makedirs("/tmp/nbldpc_sweep", exist_ok=True)
paths, cons = write_code_files("/tmp/nbldpc_sweep")
code = capi.Code.from_file(paths["divsalar.UNBLDPC.512.256.GF.256"])
B = 9472
for es in (0, 1):
    dec = capi.Decoder(code, capi.METHOD_TEMS, 10, 2, 2, 1.0, 0.5, precision=capi.PREC_F32, early_stop=es, max_batch=B)
    for i in range(2):
        c = dec.simulate(2.5, B, seed=7, first_frame=i * B, random_msg=1, crc_len=8)
    ms, n = dec.last_timing()
    it = dec.iter_timing()
    print("early_stop", es, "total %.2f ms, %d launches, iteration kernels %.2f ms:" % (ms, n, sum(it)), " ".join("%.2f" % t for t in it), "mean iterations %.2f" % (c[6] / c[0]))
