import sys, os, json, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
import numpy as np, torch
from nbldpc_b200 import capi
from config_sweep import write_code_files
os.makedirs("/tmp/nbldpc_sweep", exist_ok=True)
paths, cons = write_code_files("/tmp/nbldpc_sweep")
code = capi.Code.from_file(paths["divsalar.UNBLDPC.512.256.GF.256"])
for (nr, nc, B) in ((2, 2, 9472), (2, 3, 2368), (3, 3, 2368), (3, 2, 9472)):
    dec = capi.Decoder(code, capi.METHOD_TEMS, 10, nr, nc, 1.0, 0.5, precision=capi.PREC_F32, early_stop=0, max_batch=B)
    dec.simulate(2.5, B, seed=7, random_msg=1)
    torch.cuda.synchronize(); t0 = time.time()
    c = dec.simulate(2.5, 3 * B, seed=7, first_frame=B, random_msg=1)
    torch.cuda.synchronize(); dt = time.time() - t0
    print("n_r %d n_c %d: %.2f Mbit/s  FER %.4f (%d frames)" % (nr, nc, 3 * B * 256 / dt / 1e6, c[1] / c[0], c[0]), flush=True)
