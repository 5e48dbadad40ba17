import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tools"))
import numpy as np, torch
from nbldpc_b200 import capi
from config_sweep import write_code_files
which = sys.argv[1]
os.makedirs("/tmp/nbldpc_sweep", exist_ok=True)
paths, cons = write_code_files("/tmp/nbldpc_sweep")
if which == "cfg2":
    code = capi.Code.from_file(paths["BDS.576.288.GF.64"]); args = (capi.METHOD_EMS, 10, 8, 3, 1.0, 0.5); B = 148 * 64; kw = {}; snr = 2.0
elif which == "cfg4":
    code = capi.Code.from_file(paths["divsalar.CNBLDPC.512.256.GF.256"]); args = (capi.METHOD_QSPA, 10, 0, 0, 1.0, 0.0); B = 148 * 8
    kw = dict(modulation=256, constellation=np.array(cons["GRAY_256QAM"], np.float64)); snr = 14.0
elif which == "cfg5":
    code = capi.Code.from_hbit(paths["H_bit_GF256"], 256); args = (capi.METHOD_TEMS, 10, 2, 2, 1.0, 0.5); B = 148 * 512; kw = {}; snr = 4.0
else:
    code = capi.Code.from_file(paths["divsalar.UNBLDPC.128.64.GF.16"]); args = (capi.METHOD_QSPA, 10, 0, 0, 1.0, 0.0); B = 148 * 256; kw = {}; snr = 3.0
dec = capi.Decoder(code, *args, precision=capi.PREC_F32, early_stop=1, max_batch=B)
for i in range(3):
    c = dec.simulate(snr, B, seed=7, first_frame=i * B, random_msg=1, **kw)
print(which, c)
