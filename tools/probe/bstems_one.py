import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from nbldpc_b200 import capi
from config_sweep import write_code_files
os.makedirs("/tmp/nbldpc_sweep", exist_ok=True)
paths, cons = write_code_files("/tmp/nbldpc_sweep")
code = capi.Code.from_file(paths["divsalar.UNBLDPC.512.256.GF.256"])
B = 148 * 16
dec = capi.Decoder(code, capi.METHOD_BSTEMS, 10, 4, 3, 1.0, 0.2, precision=capi.PREC_F32, early_stop=0, max_batch=B)
for i in range(2):
    c = dec.simulate(2.75, B, seed=7, first_frame=i * B, random_msg=1)
print(c, dec.last_timing())
