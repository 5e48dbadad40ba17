"""Throughput of the decoder / field-size combinations outside the BASELINE configurations (development aid):
fp32, early stop as in the reference, device Monte-Carlo with random code words, one operating point each."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
from nbldpc_b200 import capi
from config_sweep import write_code_files
os.makedirs("/tmp/nbldpc_sweep", exist_ok=True)
paths, cons = write_code_files("/tmp/nbldpc_sweep")
G16, G64, G256 = "divsalar.UNBLDPC.128.64.GF.16", "BDS.576.288.GF.64", "divsalar.UNBLDPC.512.256.GF.256"
CASES = [("T-EMS 2/2/0.5 GF(16) (128,64)", G16, capi.METHOD_TEMS, 2, 2, 0.5, 3.0, 148 * 256),
         ("EMS 8/3/0.5 GF(16) (128,64)", G16, capi.METHOD_EMS, 8, 3, 0.5, 3.0, 148 * 256),
         ("BS-TEMS 4/3/0.2 GF(16) (128,64)", G16, capi.METHOD_BSTEMS, 4, 3, 0.2, 3.0, 148 * 256),
         ("T-EMS 2/2/0.5 GF(64) (576,288)", G64, capi.METHOD_TEMS, 2, 2, 0.5, 2.0, 148 * 64),
         ("QSPA GF(64) (576,288)", G64, capi.METHOD_QSPA, 0, 0, 0.0, 2.0, 148 * 64),
         ("BS-TEMS 4/3/0.2 GF(256) (512,256)", G256, capi.METHOD_BSTEMS, 4, 3, 0.2, 2.75, 148 * 32),
         ("EMS 8/3/0.5 GF(256) (512,256)", G256, capi.METHOD_EMS, 8, 3, 0.5, 2.75, 148 * 32),
         ("QSPA GF(256) (512,256) BPSK", G256, capi.METHOD_QSPA, 0, 0, 0.0, 2.5, 148 * 16)]
for name, cf, method, p0, p1, off, snr, B in CASES:
    code = capi.Code.from_file(paths[cf])
    dec = capi.Decoder(code, method, 10, p0, p1, 1.0, off, precision=capi.PREC_F32, early_stop=1, max_batch=B)
    dec.simulate(snr, B, seed=7, random_msg=1)
    torch.cuda.synchronize(); t0 = time.time()
    c = dec.simulate(snr, 3 * B, seed=7, first_frame=B, random_msg=1)
    torch.cuda.synchronize(); dt = time.time() - t0
    K = code.N - code.M
    print("%-36s %5.1f dB  %8.1f Mbit/s  FER %.2e  mean iterations %.2f" % (name, snr, int(c[0]) * K * code.p / dt / 1e6, c[1] / c[0], c[6] / c[0]), flush=True)
    del dec
